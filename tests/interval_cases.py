"""Shared body of the coverage-profile / interval-sum tests (CPU emulation and GPU).

The checker restates mergerHelper.formatIntervalAndFilter (misassemblyFixerLib/mergerHelper.py:137-163) with
np.sum over the downloaded coverage list, exactly as the reference computes it."""
import numpy as np

from metafinishersc_b200 import mergerHelper


def reference_formatIntervalAndFilter(repeatInterval, coverageData):
    dataList, locList = [], []
    initThres, lenThres, largestRange = 2, 1, 50000
    start, end = initThres, repeatInterval[0][0]
    if end - start > lenThres:
        locList.append([start, end])
        end = min(start + largestRange, end)
        dataList.append([np.sum(coverageData[start:end]), end - start])
    for i in range(0, len(repeatInterval) - 1):
        start, end = repeatInterval[i][1], repeatInterval[i + 1][0]
        if end - start > lenThres:
            locList.append([start, end])
            end = min(start + largestRange, end)
            dataList.append([np.sum(coverageData[start:end]), end - start])
    start, end = repeatInterval[-1][1], len(coverageData) - initThres
    if end - start > lenThres:
        locList.append([start, end])
        end = min(start + largestRange, end)
        dataList.append([np.sum(coverageData[start:end]), end - start])
    return dataList, locList


def run(E, oracle):
    rng = np.random.default_rng(5)
    lens = np.array([180_000, 1, 70_001, 33, 4096, 512, 513, 3], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    n = 30_000
    ref = rng.integers(0, len(lens), size=n).astype(np.int32)
    s1 = (rng.integers(0, 1 << 30, size=n) % lens[ref] + 1).astype(np.int32)
    e1 = (s1 - 1 + rng.integers(0, 1 << 30, size=n) % (lens[ref] - s1 + 2)).astype(np.int32)
    prof = E.coverage_profile(ref, s1, e1, off)
    want = oracle.coverage(ref, s1, e1, off)
    assert np.array_equal(prof.array(), want)
    for c in range(len(lens)):
        assert np.array_equal(prof.array(c), want[off[c]:off[c + 1]])
    # random slices, Python semantics: negative and out-of-range indices, empty and reversed slices, tile edges
    q = 4000
    qc = rng.integers(0, len(lens), size=q).astype(np.int32)
    a = rng.integers(-20, 200_000, size=q).astype(np.int64)
    b = rng.integers(-20, 200_000, size=q).astype(np.int64)
    a[:600] = a[:600] % (lens[qc[:600]] + 1)
    b[:600] = np.minimum(lens[qc[:600]], a[:600] + rng.integers(0, 1100, size=600))
    a[600:700] = (a[600:700] // 512) * 512
    b[700:800] = (b[700:800] // 512) * 512
    got = prof.sums(qc, a, b)
    for k in range(q):
        lst = want[off[qc[k]]:off[qc[k] + 1]]
        assert got[k] == int(np.sum(lst[int(a[k]):int(b[k])], dtype=np.int64)), (k, qc[k], a[k], b[k])
    assert len(prof.sums(qc[:0], a[:0], b[:0])) == 0
    # the consumer itself, against the reference's arithmetic on the downloaded lists
    names = ["Segkk%d" % i for i in range(len(lens))]
    repeatIntervalDic = {"Segkk0": [[1000, 9000], [9001, 9002], [60_000, 61_500], [179_990, 179_999]],
                         "Segkk2": [[0, 10], [20_000, 30_000]], "Segkk4": [[2, 4094]], "Segkk5": [[100, 200]], "Segkk7": [[1, 2]]}
    xDic, locDic = mergerHelper.groupCTestData(repeatIntervalDic, prof, names)
    for name, iv in repeatIntervalDic.items():
        c = names.index(name)
        wd, wl = reference_formatIntervalAndFilter(iv, want[off[c]:off[c + 1]].tolist())
        assert locDic[name] == wl
        assert xDic[name] == [[int(x[0]), x[1]] for x in wd]
    prof.free()
