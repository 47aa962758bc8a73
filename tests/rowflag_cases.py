"""Shared body of the row-predicate tests (CPU emulation and GPU): mfsc_row_flags against the reference's predicates
restated on 9-field rows exactly as the reference writes them (file:line in metafinishersc_b200/csrc/mfsc_rowflags.cuh)."""
import numpy as np

from metafinishersc_b200 import _lib, alignerRobot, engine, rowFilters


def headTailMatch(eachitem, lenDic):                     # abunHouseKeeper.py:177-202
    start1, end1, start2, end2 = eachitem[0], eachitem[1], eachitem[2], eachitem[3]
    l1, l2 = lenDic[eachitem[-2]], lenDic[eachitem[-1]]
    thres = 10
    diffContig = eachitem[-2] != eachitem[-1]
    forwardStrand = not (start2 > end2)
    headtailoverlap = (start1 <= thres and end2 >= l2 - thres) or (end1 >= l1 - thres and start2 <= thres)
    return diffContig and forwardStrand and headtailoverlap


def identicalItem(eachitem, lenDic):                     # abunHouseKeeper.py:162-175
    start1, end1, start2, end2 = eachitem[0], eachitem[1], eachitem[2], eachitem[3]
    thres = 20
    endPt1 = lenDic[eachitem[-2]] - end1
    endPt2 = lenDic[eachitem[-1]] - end2
    return abs(start1 - start2) < thres and abs(endPt1 - endPt2) < thres


def checkSatisfy(eachitem, lenDic):                      # associatedReadFinder.py:16-43
    threshold = 50
    lenSeed, lenDetect = lenDic[eachitem[-1]], lenDic[eachitem[-2]]
    checkSeed = eachitem[0] < threshold or eachitem[1] > lenSeed - threshold
    checkDetect = min(eachitem[2], eachitem[3]) < threshold or max(eachitem[2], eachitem[3]) > lenDetect - threshold
    return checkSeed and checkDetect


def completelyEmbedSR2TR(eachsub, lenDic):               # merger.py:113-123
    thres = 20
    rdStart, rdEnd = min(eachsub[2], eachsub[3]), max(eachsub[2], eachsub[3])
    return rdStart < thres and rdEnd >= lenDic[eachsub[-1]] - thres


def run(E):
    rng = np.random.default_rng(11)
    n_ref, n_qry, n = 40, 60, 20_000
    # names shared between the two sides for some records (self-alignments compare name strings)
    ref_names = ["Contig%d" % i for i in range(n_ref)]
    qry_names = ["Contig%d" % i for i in range(0, 30)] + ["Read%d" % i for i in range(30, n_qry)]
    ref_lens = rng.integers(1, 3000, size=n_ref).astype(np.int32)
    qry_lens = np.array([ref_lens[i] if i < 30 else rng.integers(1, 3000) for i in range(n_qry)], dtype=np.int32)
    r = engine.Rows()
    r.ref_id = rng.integers(0, n_ref, size=n).astype(np.int32)
    r.qry_id = rng.integers(0, n_qry, size=n).astype(np.int32)
    # coordinates concentrated near the thresholds (record ends, 10/20/50/300/390 from either end)
    def near(lens):
        edge = rng.integers(0, 2, size=n)
        d = rng.choice([0, 1, 9, 10, 11, 19, 20, 21, 49, 50, 51, 299, 300, 301, 389, 390, 391], size=n) + rng.integers(0, 2, size=n)
        v = np.where(edge == 0, 1 + d, lens - d)
        return np.clip(v, 1, np.maximum(lens, 1)).astype(np.int32)
    a, b = near(ref_lens[r.ref_id]), near(ref_lens[r.ref_id])
    r.s1, r.e1 = np.minimum(a, b), np.maximum(a, b)
    r.s2, r.e2 = near(qry_lens[r.qry_id]), near(qry_lens[r.qry_id])
    r.errors = np.zeros(n, dtype=np.int32)
    r.cols = np.ones(n, dtype=np.int32)
    rs = alignerRobot.RowSet(r, ref_names, ref_lens.tolist(), qry_names, qry_lens.tolist(), "ref.fasta", "qry.fasta")
    flags = rowFilters.row_flags(rs, E)
    lenDic = {nm: int(l) for nm, l in zip(ref_names, ref_lens)}
    lenDic.update({nm: int(l) for nm, l in zip(qry_names, qry_lens)})
    for k in range(n):
        row = [int(r.s1[k]), int(r.e1[k]), int(r.s2[k]), int(r.e2[k]), abs(int(r.e1[k]) - int(r.s1[k])) + 1,
               abs(int(r.e2[k]) - int(r.s2[k])) + 1, 100.0, ref_names[r.ref_id[k]], qry_names[r.qry_id[k]]]
        l1, l2 = lenDic[row[-2]], lenDic[row[-1]]
        want = 0
        want |= _lib.RF_HEADTAIL if headTailMatch(row, lenDic) else 0
        want |= _lib.RF_IDENTICAL if identicalItem(row, lenDic) else 0
        want |= _lib.RF_ASSOCIATED if checkSatisfy(row, lenDic) else 0
        want |= _lib.RF_EMBEDDED_LR if completelyEmbedSR2TR(row, lenDic) else 0
        want |= _lib.RF_REF_COVERED if row[4] > l1 - 300 else 0                      # readContigGraphFormer.py:173
        want |= _lib.RF_REF_HEAD if row[0] < 50 else 0                               # readContigGraphFormer.py:20
        want |= _lib.RF_SHORT_REF if abs(l1 - row[4]) < 10 else 0                    # nonRedundantResolver.py:106
        want |= _lib.RF_SHORT_QRY if abs(l2 - row[5]) < 10 else 0
        want |= _lib.RF_DIFF_NAME if row[7] != row[8] else 0
        want |= _lib.RF_END_SPAN if (row[0] < 10 and row[1] > 400 - 10) else 0        # gapFiller.py:116-118,177
        want |= _lib.RF_OVERLAP if (row[1] > l1 - 10 and row[2] < 10) else 0          # IORobot.py:529
        assert int(flags[k]) == want, (k, row, l1, l2, int(flags[k]), want)
    # the filter functions return the reference's lists
    dl = rs.datalist()
    assert rowFilters.filterData(rs, E) == [x for x in dl if headTailMatch(x, lenDic)]
    assert rowFilters.filterDataIdentical(rs, E) == [x for x in dl if not identicalItem(x, lenDic)]
    assert rowFilters.associatedRows(rs, E) == [x for x in dl if checkSatisfy(x, lenDic)]
    assert rowFilters.embeddedRows(rs, E) == [x for x in dl if completelyEmbedSR2TR(x, lenDic)]
    empty = engine.Rows()
    for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
        setattr(empty, f, np.zeros(0, dtype=np.int32))
    assert len(E.row_flags(empty, ref_lens, qry_lens, np.arange(n_ref), np.arange(n_qry))) == 0
