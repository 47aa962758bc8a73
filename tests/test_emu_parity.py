"""CPU-side parity: the kernel sources compiled for the host (one lane per warp, tests/emu) against
the oracle -- seeds, clusters, rows and DP cell counts bit-exact.  Checks the stage logic and the
host orchestration without a GPU; the GPU parity proper is tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests import cases

CASES = cases.build_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emu_matches_oracle(case, emu_engine, oracle):
    name, ref, qry, opts = case
    bad, r, o = cases.run_case(emu_engine, oracle, ref, qry, opts)
    assert not bad, "%s: %s" % (name, "; ".join(bad))
    # stage 4a (mfsc_gaps.cuh): by rule only on batches with >= 1000 matches per read, else when forced
    if name in ("many_identical_records", "gap_stage_forced_noisy", "gap_stage_forced_accurate"):
        assert r.stats["gap_tasks"] > 50 and 0 < r.stats["gap_hits"] <= r.stats["gap_tasks"]
    if name in ("indel15_reads", "reads_vs_contigs", "gap_stage_off_many_records"):
        assert r.stats["gap_hits"] == 0


def test_emu_coverage(emu_engine, oracle):
    rng = np.random.default_rng(0)
    lens = np.array([5000, 1, 70000, 33, 4096], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    n = 4000
    ref = rng.integers(0, len(lens), size=n).astype(np.int32)
    s1 = (rng.integers(0, 1 << 30, size=n) % lens[ref] + 1).astype(np.int32)
    e1 = (s1 - 1 + rng.integers(0, 1 << 30, size=n) % (lens[ref] - s1 + 2)).astype(np.int32)
    e1[::7] = lens[ref][::7]
    cov = emu_engine.coverage(ref, s1, e1, off)
    assert np.array_equal(cov, oracle.coverage(ref, s1, e1, off))
    assert np.array_equal(emu_engine.coverage(ref[:0], s1[:0], e1[:0], off), np.zeros(off[-1], dtype=np.int32))


def test_emu_coverage_profile_interval_sums(emu_engine, oracle):
    from tests import interval_cases
    interval_cases.run(emu_engine, oracle)


def test_emu_row_flags(emu_engine):
    from tests import rowflag_cases
    rowflag_cases.run(emu_engine)
