"""Seeded parity cases shared by the emulation (CPU) and GPU parity tests.

Each case is (name, reference records, query records, nucmer options).  Shapes follow the call
sites of SURVEY.md Appendix A: reads vs contigs (#4/#10), contig self-alignment in --maxmatch
(#2/#3/#5) and default mode (#1), doubled reads vs doubled reads (#8), small -l 10 pairs (#12),
contig ends vs reads (#6)."""
import numpy as np

from metafinishersc_b200 import datagen

_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def rc(s):
    return s.translate(_COMP)[::-1]


def _mutate(s, i, rng):
    b = bytearray(s)
    for p in rng.integers(0, len(b), size=5):
        b[p] = ord("N")
    return bytes(b).lower() if i % 3 == 0 else bytes(b)


def build_cases(scale=1):
    rng = np.random.default_rng(5)
    d = datagen.abun_gap(G=200_000 * scale, repeat_len=1200, L=3000, seed=1)
    reads = d["reads"]
    mixed = [rc(r) if i % 2 else r for i, r in enumerate(reads[:160 * scale])]
    cases = []
    cases.append(("reads_vs_contigs", d["contigs"], reads[:200 * scale], {}))
    cases.append(("both_strands", d["contigs"], mixed, {}))
    cases.append(("reference_self", d["reference"], d["reference"], {}))
    cases.append(("contigs_vs_reference", d["reference"], d["contigs"], {}))
    weird = [_mutate(r, i, rng) for i, r in enumerate(mixed[:60])] + \
        [b"", b"ACGT", b"A" * 25, reads[0][:19], reads[1][:20], reads[2][:64], reads[3][:65], reads[4][:66]]
    cases.append(("ragged_and_N_queries", d["contigs"], weird, {}))
    cases.append(("N_in_reference", [_mutate(c, 1, rng) for c in d["contigs"]], mixed[:60], {}))
    dbl = []
    for r in reads[:80]:
        dbl += [r, rc(r)]
    cases.append(("doubled_reads_all_vs_all", dbl, dbl[:60], {}))
    cases.append(("small_pair_l10", [d["contigs"][0][1000:9000]], [d["contigs"][0][7000:15000]], dict(minmatch=10)))
    cases.append(("default_mode_self", d["contigs"], d["contigs"], dict(maxmatch=False)))
    cases.append(("default_mode_reads", d["contigs"], mixed[:80], dict(maxmatch=False)))
    cases.append(("breaklen_50", d["contigs"], mixed[:80], dict(breaklen=50)))
    cases.append(("nosimplify", d["contigs"], mixed[:80], dict(simplify=False)))
    trunc = []
    for c in d["contigs"]:
        trunc += [c[:400], c[-400:], rc(c[:400]), rc(c[-400:])]
    cases.append(("contig_ends_vs_reads", trunc, mixed[:120], {}))
    d2 = datagen.sc_lr(G=300_000, n_reads=12 * scale, seed=2)
    cases.append(("indel15_reads", d2["contigs"], [rc(r) if i % 2 else r for i, r in enumerate(d2["reads"])], {}))
    # tandem repeats: large match components for the chaining DP
    unit = datagen.random_genome(37, np.random.default_rng(9)).tobytes()
    flank = datagen.random_genome(4000, np.random.default_rng(10)).tobytes()
    tandem = flank[:2000] + unit * 40 + flank[2000:]
    cases.append(("tandem_repeat", [tandem], [flank[1500:2000] + unit * 25 + flank[2000:2400], rc(tandem[1000:3500])], {}))
    # the DP band limit (band_cap) actually binding, and extensions longer than the 10 000-base DP limit
    cases.append(("band_cap_64", d2["contigs"], d2["reads"][:6], dict(band_cap=64)))
    long_reads, _ = datagen.noisy_reads(np.frombuffer(d["contigs"][0], dtype=np.uint8), 6, 26_000, 0.01, 0.01, np.random.default_rng(31))
    cases.append(("reads_longer_than_dp_limit", d["contigs"], [long_reads[0], rc(long_reads[1])] + long_reads[2:], {}))
    # 2000 identical reference records: every read has 2000 clusters, and the first match buffer overflows (retry path)
    unit300 = datagen.random_genome(300, np.random.default_rng(12)).tobytes()
    many = [unit300] * 2000
    cases.append(("many_identical_records", many, [unit300, rc(unit300), unit300[10:280]] * 14, {}))
    cases.append(("explicit_kmer_8", d["contigs"], mixed[:40], dict(kmer=8)))
    cases.append(("explicit_kmer_14", d["contigs"], mixed[:40], dict(kmer=14)))
    # DP engines: the general engine alone, and the packed engine with field spans so small that some / all reads are
    # handed over to the general engine
    cases.append(("engine_general_only", d2["contigs"], d2["reads"][:10], dict(engine=1)))
    # stage 4a (forward extensions as tasks ahead of the walk; by rule only with >= 1000 matches per read, as in
    # many_identical_records above): forced on noisy and on accurate reads, forced off where the rule turns it on
    cases.append(("gap_stage_forced_noisy", d2["contigs"], [rc(r) if i % 2 else r for i, r in enumerate(d2["reads"])], dict(engine=2)))
    cases.append(("gap_stage_forced_accurate", d["contigs"], mixed[:80], dict(engine=2)))
    cases.append(("gap_stage_off_many_records", many, [unit300, rc(unit300), unit300[10:280]] * 6, dict(engine=3)))
    cases.append(("packed_some_reads_handed_over", d2["contigs"], d2["reads"][:24], dict(pk_m_span=16)))
    cases.append(("packed_all_reads_handed_over", d2["contigs"], d2["reads"][:8], dict(pk_u_span=100)))
    return cases


def run_case(E, O, ref, qry, opts):
    """Run one case through engine E and the oracle O; returns a list of mismatch descriptions."""
    from metafinishersc_b200 import engine
    okw = {k: v for k, v in opts.items() if k not in ("kmer", "engine", "pk_u_span", "pk_m_span")}
    okw.setdefault("band_cap", 248)
    o = O.align(ref, qry, **okw)
    P = engine.make_params(**opts)
    ix = E.index(ref, P)
    r = E.align(ix, qry, P, flags=engine.KEEP_MEMS | engine.KEEP_CLUSTERS)
    ix.free()
    bad = []
    if not np.array_equal(o["mems"], r.mems):
        bad.append("seeds differ: oracle %s, library %s" % (o["mems"].shape, r.mems.shape))
    if not np.array_equal(o["clusters"][1], r.clusters[1]):
        bad.append("cluster matches differ: oracle %s, library %s" % (o["clusters"][1].shape, r.clusters[1].shape))
    got = np.stack([r.ref_id, r.qry_id, (r.s2 > r.e2).astype(np.int32), r.s1, r.e1, r.s2, r.e2, r.errors, r.cols],
                   axis=1).astype(np.int64) if len(r) else np.zeros((0, 9), dtype=np.int64)
    want = o["rows"][:, :9]
    if not np.array_equal(got, want):
        n = min(len(got), len(want))
        first = next((i for i in range(n) if not np.array_equal(got[i], want[i])), n)
        bad.append("rows differ (%d vs %d), first at %d: %s vs %s" % (len(got), len(want), first,
                   got[first].tolist() if first < len(got) else None, want[first].tolist() if first < len(want) else None))
    if len(r) and r.stats["cells"] != o["stats"]["cells"]:
        bad.append("DP cells differ: %d vs %d" % (r.stats["cells"], o["stats"]["cells"]))
    if len(r) and (r.stats["cap_hits"] != o["stats"]["cap_hits"] or r.stats["max_band"] != o["stats"]["max_band"]):
        bad.append("band statistics differ: cap hits %d vs %d, widest band %d vs %d" % (r.stats["cap_hits"], o["stats"]["cap_hits"],
                                                                                      r.stats["max_band"], o["stats"]["max_band"]))
    # delta mode: same rows, plus the indel offset list of every row bit-exact
    ix = E.index(ref, P)
    rd = E.align(ix, qry, P, flags=engine.WANT_DELTAS)
    ix.free()
    gotd = np.stack([rd.ref_id, rd.qry_id, (rd.s2 > rd.e2).astype(np.int32), rd.s1, rd.e1, rd.s2, rd.e2, rd.errors, rd.cols],
                    axis=1).astype(np.int64) if len(rd) else np.zeros((0, 9), dtype=np.int64)
    if not np.array_equal(gotd, want):
        bad.append("delta mode changes the rows")
    else:
        for i in range(len(rd)):
            mine = rd.deltas[rd.delta_begin[i]:rd.delta_end[i]]
            ref_d = o["deltas"][o["rows"][i, 9]:o["rows"][i, 10]]
            if rd.delta_begin[i] < 0 or not np.array_equal(mine.astype(np.int64), ref_d):
                bad.append("delta list of row %d differs: %s... vs %s..." % (i, mine[:6].tolist(), ref_d[:6].tolist()))
                break
    return bad, r, o
