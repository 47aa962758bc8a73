"""GPU runs of the host mirror (alignerRobot / IORobot / coverage consumers) and of the full-size
known answers and properties."""
import json
import os

import numpy as np
import pytest

from metafinishersc_b200 import datagen, engine
from tests import test_golden, test_host_pipeline

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_host_pipeline_gpu(gpu_engine, oracle, tmp_path):
    test_host_pipeline.run_pipeline(gpu_engine, oracle, tmp_path)


def rows_in_delta_order(r, rnames, qnames):
    idy = r.idy()
    return [[int(r.s1[i]), int(r.e1[i]), int(r.s2[i]), int(r.e2[i]), abs(int(r.e1[i]) - int(r.s1[i])) + 1,
             abs(int(r.e2[i]) - int(r.s2[i])) + 1, float("%.2f" % float(idy[i])), rnames[r.ref_id[i]], qnames[r.qry_id[i]]]
            for i in range(len(r))]


def test_gpu_reproduces_readme_tables(gpu_engine):
    """Reference README.md:111-129 at full 5 Mbp scale: whole-contig exact matches and the delta order."""
    gold = json.load(open(os.path.join(GOLD, "readme_tables.json")))
    ref, contigs = test_golden.readme_genomes()
    names = ["Segkk0", "Segkk1"]
    P = engine.make_params()
    ix = gpu_engine.index(ref, P)
    assert rows_in_delta_order(gpu_engine.align(ix, ref, P), names, names) == gold["reference_vs_reference"]
    ix.free()
    ix = gpu_engine.index(contigs, P)
    assert rows_in_delta_order(gpu_engine.align(ix, ref, P), names, names) == gold["LC_vs_reference"]
    ix.free()


def test_full_size_properties(gpu_engine):
    """BASELINE cfg2 shape at a size the oracle cannot finish in seconds: size-independent properties."""
    d = datagen.sc_lr(G=5_000_000, n_reads=3000, seed=100, read_seed=2100)
    P = engine.make_params()
    ix = gpu_engine.index(d["contigs"], P)
    a = gpu_engine.align(ix, d["reads"], P)
    b = gpu_engine.align(ix, d["reads"], P)
    for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
        assert np.array_equal(getattr(a, f), getattr(b, f))                    # idempotent / deterministic
    clen = len(d["contigs"][0])
    rl = np.array([len(x) for x in d["reads"]])
    assert len(a) >= 0.9 * len(d["reads"])
    assert (a.s1 >= 1).all() and (a.e1 <= clen).all() and (a.s1 <= a.e1).all()
    lo, hi = np.minimum(a.s2, a.e2), np.maximum(a.s2, a.e2)
    assert (lo >= 1).all() and (hi <= rl[a.qry_id]).all()
    assert (a.cols >= np.maximum(a.e1 - a.s1, hi - lo) + 1).all() and (a.errors >= 0).all() and (a.errors < a.cols).all()
    assert (a.s2 < a.e2).all()                                                 # generator emits forward reads only
    # coverage: checksum of checksums -- the profile sums to the summed interval lengths, and is linear in the rows
    off = np.array([0, clen], dtype=np.int64)
    cov = gpu_engine.coverage(a.ref_id, a.s1, a.e1, off)
    assert int(cov.sum()) == int((a.e1 - a.s1 + 1).sum())
    h = len(a) // 2
    c1 = gpu_engine.coverage(a.ref_id[:h], a.s1[:h], a.e1[:h], off)
    c2 = gpu_engine.coverage(a.ref_id[h:], a.s1[h:], a.e1[h:], off)
    assert np.array_equal(c1 + c2, cov)
    # reverse-complemented reads land on the same contig interval with the strand flipped
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    sub = d["reads"][:200]
    r1 = gpu_engine.align(ix, sub, P)
    r2 = gpu_engine.align(ix, [x.translate(comp)[::-1] for x in sub], P)
    k1 = sorted(zip(r1.qry_id.tolist(), r1.s1.tolist(), r1.e1.tolist(), r1.s2.tolist(), r1.e2.tolist(), r1.errors.tolist()))
    L = np.array([len(x) for x in sub])
    k2 = sorted(zip(r2.qry_id.tolist(), r2.s1.tolist(), r2.e1.tolist(), (L[r2.qry_id] + 1 - r2.s2).tolist(),
                    (L[r2.qry_id] + 1 - r2.e2).tolist(), r2.errors.tolist()))
    assert k1 == k2
    ix.free()


def test_engines_agree_on_substitution_errors(gpu_engine):
    """Both DP engines give the same rows on reads with substitutions + indels (the reference's generator makes indels only),
    at a size the oracle does not finish in seconds; the packed engine does not hand reads over on such data."""
    rng = np.random.default_rng(7)
    d = datagen.sc_lr(G=2_000_000, n_reads=1500, seed=100, read_seed=2100, p_del=0.04, p_ins=0.04)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for r in d["reads"]:
        a = np.frombuffer(r, dtype=np.uint8).copy()
        m = rng.random(len(a)) < 0.06
        a[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        reads.append(a.tobytes())
    out = []
    for eng in (0, 1, 2, 3):        # packed (stage 4a by rule), general, packed with stage 4a forced / off
        P = engine.make_params(engine=eng)
        ix = gpu_engine.index(d["contigs"], P)
        out.append(gpu_engine.align(ix, reads, P))
        ix.free()
    for o in out[1:]:
        for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
            assert np.array_equal(getattr(out[0], f), getattr(o, f)), f
        assert out[0].stats["cells"] == o.stats["cells"] and out[0].stats["max_band"] == o.stats["max_band"]
    assert out[2].stats["gap_hits"] > 0 and out[3].stats["gap_hits"] == 0
    assert out[0].stats["dp_retry_reads"] == 0 and len(out[0]) >= 1400


@pytest.mark.parametrize("opts", [dict(breaklen=50), dict(minmatch=10), dict(band_cap=64), dict(simplify=False), dict(maxmatch=False),
                                  dict(breaklen=20, band_cap=40)],
                         ids=["b50", "l10", "cap64", "nosimplify", "mum", "b20cap40"])
def test_engines_agree_under_options(gpu_engine, opts):
    """Packed (with its small-rectangle path, whose size limit follows breaklen and band_cap) vs general engine under the
    option sets the wrapper uses and a few extreme ones: both strands, accurate and noisy reads, reads with N runs."""
    rng = np.random.default_rng(23)
    d1 = datagen.abun_gap(G=400_000, repeat_len=3000, L=4000, p=0.01, abun=(3, 2), seed=5)
    d2 = datagen.sc_lr(G=400_000, n_reads=150, seed=6, p_del=0.05, p_ins=0.05, L=5000)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for i, r in enumerate(d1["reads"][:400] + d2["reads"]):
        if i % 3 == 1:
            r = r.translate(comp)[::-1]
        if i % 11 == 0:
            k = int(rng.integers(100, len(r) - 100))
            r = r[:k] + b"N" * int(rng.integers(1, 30)) + r[k + 5:]
        reads.append(r)
    contigs = d1["contigs"] + d2["contigs"]
    out = []
    for eng in (0, 1):
        P = engine.make_params(engine=eng, **opts)
        ix = gpu_engine.index(contigs, P)
        out.append(gpu_engine.align(ix, reads, P, flags=engine.WANT_DELTAS))
        ix.free()
    a, b = out
    for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for i in range(len(a)):        # the slices of the shared delta buffer are handed out in completion order: compare the lists
        assert np.array_equal(a.deltas[a.delta_begin[i]:a.delta_end[i]], b.deltas[b.delta_begin[i]:b.delta_end[i]]), i
    assert a.stats["cells"] == b.stats["cells"] and a.stats["dp_calls"] == b.stats["dp_calls"] and a.stats["max_band"] == b.stats["max_band"]
    assert len(a) > 300 and a.stats["dp_retry_reads"] > 0          # the reads with N went through the general engine


def test_cfg5_scale_index(gpu_engine):
    """BASELINE cfg 5 shape on the reference side: 100 contigs x 5 Mbp (500 Mbp index, k = 15), one read batch
    drawn from known places.  Property: every read's longest row lands on its source contig and offset."""
    rng = np.random.default_rng(77)
    n_contigs, G, L, n_reads = 100, 5_000_000, 10_000, 4000
    contigs = [datagen.random_genome(G, np.random.default_rng(5000 + c)) for c in range(n_contigs)]
    src = rng.integers(0, n_contigs, size=n_reads)
    reads, starts = [], []
    for c in range(n_contigs):
        idx = np.nonzero(src == c)[0]
        r, st = datagen.noisy_reads(contigs[c], len(idx), L, 0.01, 0.01, np.random.default_rng(9000 + c))
        reads += r
        starts += [(c, int(s)) for s in st]
    P = engine.make_params()
    buf = np.concatenate(contigs)
    off = np.arange(n_contigs + 1, dtype=np.int64) * G
    ix = gpu_engine.index((buf, off), P)
    info = ix.info()
    assert info["k"] == 15 and info["bases"] == n_contigs * G
    r = gpu_engine.align(ix, reads, P)
    ix.free()
    best = {}
    for i in range(len(r)):
        q = int(r.qry_id[i])
        ln = int(r.e1[i] - r.s1[i])
        if q not in best or ln > best[q][0]:
            best[q] = (ln, int(r.ref_id[i]), int(r.s1[i]))
    assert len(best) == n_reads
    ok = sum(1 for q, (ln, c, s1) in best.items() if c == starts[q][0] and abs(s1 - 1 - starts[q][1]) <= 50 and ln > 9000)
    assert ok == n_reads


def test_batch_over_two_gpus(gpu_engine, oracle, tmp_path):
    """In-process fan-out of the 20 read parts over two GPUs (one host thread per GPU): same files as the serial loop."""
    from metafinishersc_b200 import alignerRobot, houseKeeper
    if gpu_engine.device_count() < 2:
        pytest.skip("needs two GPUs")
    alignerRobot.set_engine(gpu_engine)
    folder, d, reads = test_host_pipeline.make_folder(tmp_path, n_reads=200)
    serial = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", "serial")
    houseKeeper.globalGPUs = 2
    try:
        par = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", "par")
    finally:
        houseKeeper.globalGPUs = 1
    assert par == serial and len(par) > 0
    for i in range(1, 21):
        idx = alignerRobot.zeropadding(i)
        assert open(folder + "par" + idx + ".delta").read().split("\n")[1:] == open(folder + "serial" + idx + ".delta").read().split("\n")[1:]
    alignerRobot.set_engine(None)


def test_pipelined_align_equals_one_call_gpu(gpu_engine):
    """Engine.align in pieces (upload thread + two aligner threads on their own streams) returns the rows, statistics and delta
    lists of one call."""
    d = datagen.abun_gap(G=300_000, repeat_len=2000, L=4000, p=0.01, abun=(6, 4), seed=21)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = [r.translate(comp)[::-1] if i % 2 else r for i, r in enumerate(d["reads"])] + [b"", b"ACGT"]
    P = engine.make_params()
    ix = gpu_engine.index(d["contigs"], P)
    for flags in (0, engine.WANT_DELTAS):
        one = gpu_engine.align(ix, reads, P, flags=flags, pieces=1)
        for k in (2, 5):
            many = gpu_engine.align(ix, reads, P, flags=flags, pieces=k)
            for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
                assert np.array_equal(getattr(one, f), getattr(many, f)), (f, k)
            assert one.stats["cells"] == many.stats["cells"] and one.stats["max_band"] == many.stats["max_band"] and len(one) > 500
            if flags:
                for i in range(0, len(one), 7):
                    assert np.array_equal(one.deltas[one.delta_begin[i]:one.delta_end[i]], many.deltas[many.delta_begin[i]:many.delta_end[i]])
    ix.free()


def test_gap_stage_on_reads_vs_reads(gpu_engine):
    """Reads against reads at high depth (BASELINE config 4 shape: > 1000 matches per read): stage 4a switches itself on, and the
    rows, cell counts and widest band are those of the plain walk (engine 3)."""
    g = datagen.random_genome(150_000, np.random.default_rng(77))
    reads, _ = datagen.noisy_reads(g, 700, 6000, 0.01, 0.01, np.random.default_rng(78))
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    dbl = []
    for r in reads:
        dbl += [r, r.translate(comp)[::-1]]
    out = []
    for eng in (0, 3):
        P = engine.make_params(engine=eng)
        ix = gpu_engine.index(dbl, P)
        out.append(gpu_engine.align(ix, dbl, P))
        ix.free()
    a, b = out
    assert a.stats["mems"] >= 1000 * len(dbl) and a.stats["gap_hits"] > 10_000 and b.stats["gap_hits"] == 0
    for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.stats["cells"] == b.stats["cells"] and a.stats["max_band"] == b.stats["max_band"] and a.stats["dp_calls"] == b.stats["dp_calls"]
