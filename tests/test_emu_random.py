"""Randomised parity on the CPU: the kernel sources compiled for the host (tests/emu) against the oracle on seeded random
scenarios -- genomes with tandem and inverted repeats, reads with substitutions, indels, N runs and either strand, and
random option sets (minmatch, breaklen, band cap, maxmatch, simplify).  Seeds, clusters, rows, DP cell counts and the
delta lists must be bit-exact, in rows-only and in delta mode (tests/cases.run_case)."""
import numpy as np
import pytest

from tests import cases

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def scenario(seed):
    rng = np.random.default_rng(1000 + seed)
    G = int(rng.integers(8_000, 40_000))
    g = ACGT[rng.integers(0, 4, size=G, dtype=np.uint8)].copy()
    if rng.random() < 0.6:                       # tandem repeat
        u = int(rng.integers(5, 60)); n = int(rng.integers(3, 30)); at = int(rng.integers(0, G - u * n - 1))
        g[at:at + u * n] = np.tile(g[at:at + u], n)
    if rng.random() < 0.6:                       # dispersed copy, sometimes inverted
        ln = int(rng.integers(200, 2500)); a = int(rng.integers(0, G - ln)); b = int(rng.integers(0, G - ln))
        seg = g[a:a + ln].tobytes()
        if rng.random() < 0.5:
            seg = seg.translate(COMP)[::-1]
        g[b:b + ln] = np.frombuffer(seg, dtype=np.uint8)
    n_contigs = int(rng.integers(1, 4))
    cuts = sorted(rng.integers(1000, G - 1000, size=n_contigs - 1).tolist())
    contigs = [g[a:b].tobytes() for a, b in zip([0] + cuts, cuts + [G])]
    if rng.random() < 0.3:                       # N run inside a contig
        c = contigs[0]; k = int(rng.integers(10, len(c) - 40))
        contigs[0] = c[:k] + b"N" * int(rng.integers(1, 25)) + c[k + 3:]
    reads = []
    for _ in range(int(rng.integers(4, 14))):
        L = int(rng.integers(150, 5000)); s = int(rng.integers(0, G - L))
        r = g[s:s + L].copy()
        psub, pdel, pins = rng.choice([0.0, 0.01, 0.05]), rng.choice([0.0, 0.01, 0.06]), rng.choice([0.0, 0.01, 0.06])
        m = rng.random(L) < psub
        r[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
        keep = rng.random(L) >= pdel
        ins = rng.random(L) < pins
        both = np.empty(2 * L, dtype=np.uint8); both[0::2] = r; both[1::2] = ACGT[rng.integers(0, 4, size=L)]
        mask = np.empty(2 * L, dtype=bool); mask[0::2] = keep; mask[1::2] = ins
        x = both[mask].tobytes()
        if rng.random() < 0.5:
            x = x.translate(COMP)[::-1]
        if rng.random() < 0.15 and len(x) > 80:
            k = int(rng.integers(20, len(x) - 40)); x = x[:k] + b"N" * int(rng.integers(1, 12)) + x[k + 2:]
        reads.append(x)
    if rng.random() < 0.3:
        reads.append(b"")                        # empty read
    opts = dict(minmatch=int(rng.choice([10, 15, 20])), breaklen=int(rng.choice([20, 50, 200])),
                band_cap=int(rng.choice([24, 64, 248])), maxmatch=bool(rng.random() < 0.7), simplify=bool(rng.random() < 0.8))
    if rng.random() < 0.3:
        opts["pk_m_span"] = int(rng.integers(1, 20))     # push some reads through the hand-over path
    opts["engine"] = int(rng.choice([0, 2, 2, 3]))       # stage 4a by rule / forced / off (same rows in every mode)
    return contigs, reads, opts


@pytest.mark.parametrize("seed", range(36))
def test_emu_random_scenario(seed, emu_engine, oracle):
    contigs, reads, opts = scenario(seed)
    bad, r, o = cases.run_case(emu_engine, oracle, contigs, reads, opts)
    assert not bad, "seed %d %s: %s" % (seed, opts, "; ".join(bad))
