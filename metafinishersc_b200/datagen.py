"""Seeded synthetic data of the shapes BASELINE.json names.

Restates the *model* of the reference's generator (srcRefactor/dataGenLib/dataGenLib.py:26-33
uniform ACGT genome, :64-77 repeat insertion, :181-204 indel-only noisy forward-strand
reads; dataGenerator.py:218-272 abunGap_example, :144-186 SC_LR_example) with numpy
Generators instead of Python 2's `random` stream (which the reference never seeds and
which is not reproducible under Python 3).  Fixture generator, not part of the hot path.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_genome(G, rng):
    return _ACGT[rng.integers(0, 4, size=G, dtype=np.uint8)]


def noisy_reads(genome, n_reads, L, p_del, p_ins, rng):
    """dataGenLib.createANoisyRead (dataGenLib.py:181-204), vectorised: each template base is
    dropped with prob p_del and followed by one uniform inserted base with prob p_ins."""
    G = len(genome)
    starts = rng.integers(0, G - L + 1, size=n_reads)
    out = []
    for s in starts:
        seg = genome[s:s + L]
        keep = rng.random(L) >= p_del
        ins = rng.random(L) < p_ins
        ins_base = _ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        # interleave: position j contributes seg[j] (if kept) then ins_base[j] (if inserted)
        both = np.empty(2 * L, dtype=np.uint8)
        both[0::2] = seg
        both[1::2] = ins_base
        mask = np.empty(2 * L, dtype=bool)
        mask[0::2] = keep
        mask[1::2] = ins
        out.append(both[mask].tobytes())
    return out, starts


def abun_gap(G=5_000_000, repeat_len=12_000, L=6000, p=0.01, abun=(50, 20), seed=0):
    """README test-set shape (dataGenerator.py:218-272): two genomes sharing one segment,
    two chimeric contigs, reads at 50X/20X.  Returns dict(reference, contigs, reads)."""
    m = len(abun)
    genomes = [random_genome(G, np.random.default_rng(1000 + seed * 100 + g)) for g in range(m)]
    start = G // 2
    for g in range(1, m):
        genomes[g][start:start + repeat_len] = genomes[0][start:start + repeat_len]
    half = G // 2
    contigs = [np.concatenate([genomes[0][:half], genomes[1][half:]]),
               np.concatenate([genomes[1][:half], genomes[0][half:]])]
    reads = []
    for g in range(m):
        n = int(G * abun[g] / L)
        r, _ = noisy_reads(genomes[g], n, L, p, p, np.random.default_rng(2000 + seed * 100 + g))
        reads += r
    return dict(reference=[x.tobytes() for x in genomes], contigs=[x.tobytes() for x in contigs], reads=reads)


def sc_lr(G=5_000_000, L=10_000, cov=50, p_del=0.075, p_ins=0.075, seed=0, n_reads=None, read_seed=None):
    """BASELINE config 2 (SC_LR_example shape, dataGenerator.py:144-186): one genome with three
    1-kb repeat copies 1 Mbp apart (scaled with G), contig = seg0+seg2+seg1+seg3, reads with
    15% indel error (p_del = p_ins = 0.075)."""
    rng = np.random.default_rng(1000 + seed)
    genome = random_genome(G, rng)
    sep = G // 5
    rep = min(1000, sep // 4)
    for i in range(2, 4):
        genome[i * sep:i * sep + rep] = genome[sep:sep + rep]
    segs = [genome[:sep], genome[sep:2 * sep], genome[2 * sep:3 * sep], genome[3 * sep:]]
    contig = np.concatenate([segs[0], segs[2], segs[1], segs[3]])
    n = int(G * cov / L) if n_reads is None else n_reads
    reads, _ = noisy_reads(genome, n, L, p_del, p_ins, np.random.default_rng(2000 + seed if read_seed is None else read_seed))
    return dict(reference=[genome.tobytes()], contigs=[contig.tobytes()], reads=reads)


def names(prefix, n):
    """dataGenLib.writeListToFile naming (dataGenLib.py:16-24): >Segkk<i>."""
    return ["%s%d" % (prefix, i) for i in range(n)]
