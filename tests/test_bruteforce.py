"""A second opinion on the oracle's DP: oracle/bruteforce.cpp (unbanded full matrices, written independently) against
the engine of oracle/mfsc_oracle.cpp on windows of up to 2 kbp, and row-level invariants that hold for ANY correct
alignment (error columns >= unit-cost edit distance; the delta list re-scored from the sequences gives the row's
errors and columns).  The CUDA kernels are bit-exact against the oracle (tests/test_gpu_parity.py), so what is
established here for the oracle holds for them."""
import numpy as np
import pytest

from metafinishersc_b200 import datagen
from oracle import bruteforce as BF
from oracle import oracle as O

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _mutate(seq, rng, p_sub, p_ins, p_del):
    out = bytearray()
    for c in seq:
        r = rng.random()
        if r < p_del:
            continue
        if r < p_del + p_sub:
            out.append(rng.choice([x for x in b"ACGT" if x != c]))
        else:
            out.append(c)
        if rng.random() < p_ins:
            out.append(b"ACGT"[rng.integers(0, 4)])
    return bytes(out)


def _windows():
    rng = np.random.default_rng(77)
    out = []
    for n, (ps, pi, pd) in [(40, (0.05, 0.02, 0.02)), (90, (0.0, 0.08, 0.08)), (150, (0.1, 0.0, 0.0)), (300, (0.02, 0.01, 0.01)),
                            (600, (0.03, 0.06, 0.06)), (1000, (0.0, 0.075, 0.075)), (1500, (0.05, 0.05, 0.05)), (2000, (0.01, 0.01, 0.01)),
                            (2000, (0.0, 0.075, 0.075)), (700, (0.15, 0.0, 0.0)), (64, (0.0, 0.0, 0.0)), (1200, (0.0, 0.0, 0.12))]:
        a = datagen.random_genome(n, rng).tobytes()
        out.append((a, _mutate(a, rng, ps, pi, pd)))
    # gap-only cases: one long insertion / deletion, and a tandem repeat where many paths tie
    a = datagen.random_genome(500, rng).tobytes()
    out.append((a, a[:200] + datagen.random_genome(35, rng).tobytes() + a[200:]))
    out.append((a, a[:300] + a[340:]))
    unit = datagen.random_genome(7, rng).tobytes()
    out.append((a[:100] + unit * 12 + a[100:200], a[:100] + unit * 9 + a[100:200]))
    return out


@pytest.mark.parametrize("k", range(15))
def test_forced_dp_is_globally_optimal(k):
    """The DP that realises the path between two matches of a cluster (forced, target = far corner): it reaches the corner and
    the path it reports has the optimal global affine score of the unbanded matrix."""
    a, b = _windows()[k]
    r = O.dp(a, b, forced=True, optimal=False)
    assert r["reached"] and r["fi"] == len(a) and r["fj"] == len(b)
    s = BF.score_ops(a, b, r["ops"])
    assert (s["a_used"], s["b_used"]) == (len(a), len(b)) and s["cols"] == len(r["ops"])
    assert s["score"] == BF.global_score(a, b)
    assert s["errors"] >= BF.edit_distance(a, b)


@pytest.mark.parametrize("k", range(15))
def test_search_dp_end_cell(k):
    """The extending DP (not forced): junk behind the similar stretch, so the extension has to stop inside.  End cell ==
    the one an unbanded X-drop walk ends on (same break length, nothing trimmed), and the reported path scores what
    the full matrix holds in that cell."""
    a, b = _windows()[k]
    rng = np.random.default_rng(1000 + k)
    a2 = a + datagen.random_genome(400, rng).tobytes()
    b2 = b + datagen.random_genome(400, rng).tobytes()
    r = O.dp(a2, b2, forced=False, optimal=True)
    e = BF.extend(a2, b2, 200)
    assert (r["fi"], r["fj"]) == (e["fi"], e["fj"])
    s = BF.score_ops(a2, b2, r["ops"])
    assert (s["a_used"], s["b_used"]) == (r["fi"], r["fj"]) and s["score"] == e["score"]
    assert not r["reached"]


def test_short_break_length():
    a, b = _windows()[5]
    rng = np.random.default_rng(5)
    a2, b2 = a + datagen.random_genome(200, rng).tobytes(), b + datagen.random_genome(200, rng).tobytes()
    r = O.dp(a2, b2, breaklen=50, forced=False, optimal=True)
    e = BF.extend(a2, b2, 50)
    assert (r["fi"], r["fj"]) == (e["fi"], e["fj"])


def _strand(seq, d):
    return seq if d == 0 else seq.translate(_COMP)[::-1]


@pytest.mark.parametrize("shape", ["accurate", "indel15", "both_strands_subs"])
def test_rows_against_edit_distance_and_delta_rescoring(shape):
    """Every row the oracle reports: the delta list, walked over the two sequences, consumes exactly [S1,E1] x [S2,E2] and holds
    exactly `errors` non-matching columns out of `cols`; and no alignment of those two substrings can have fewer error
    columns than their unit-cost edit distance."""
    if shape == "accurate":
        d = datagen.abun_gap(G=60_000, repeat_len=800, L=2500, seed=4)
        ref, qry = d["contigs"], d["reads"][:40]
    elif shape == "indel15":
        d = datagen.sc_lr(G=80_000, L=3000, n_reads=12, seed=6)
        ref, qry = d["contigs"], d["reads"]
    else:
        d = datagen.abun_gap(G=60_000, repeat_len=800, L=2500, seed=5)
        rng = np.random.default_rng(8)
        qry = [_strand(_mutate(r, rng, 0.03, 0.0, 0.0), i & 1) for i, r in enumerate(d["reads"][:30])]
        ref = d["contigs"]
    o = O.align(ref, qry, band_cap=1 << 20)
    assert len(o["rows"]) > 0
    for row in o["rows"]:
        rid, qid, dirn, sA, eA, sB, eB, errors, cols, db, de = [int(x) for x in row]
        a = ref[rid][sA - 1:eA]
        q = _strand(qry[qid], dirn)
        L = len(qry[qid])
        s2, e2 = (sB, eB) if dirn == 0 else (L - sB + 1, L - eB + 1)
        b = q[s2 - 1:e2]
        ops, used = [], 0
        for dl in o["deltas"][db:de]:
            run = abs(int(dl)) - 1
            ops.append("M" * run + ("I" if dl > 0 else "D"))
            used += run + 1
        ops = "".join(ops)
        ops += "M" * (cols - used)
        s = BF.score_ops(a, b, ops)
        assert (s["a_used"], s["b_used"], s["cols"]) == (len(a), len(b), cols)
        assert s["errors"] == errors
        assert errors >= BF.edit_distance(a, b)
        assert cols - errors <= min(len(a), len(b))
