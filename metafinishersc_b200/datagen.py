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


# ------------------------------------------------------------------------------------------
# BASELINE configs 3-5: metagenomes.  Reads come out as (uint8 bases, int64 offsets) arrays, generated
# in independent chunks of CHUNK reads (seed = f(genome, chunk)), so that any rank of a sharded run can produce exactly
# its own slice of the global read order (genome 0's reads first, then genome 1's, ...) without the others.
# ------------------------------------------------------------------------------------------
CHUNK = 1000


def _reads_chunk(genome, n, L, p_del, p_ins, seed):
    """n noisy forward-strand reads (dataGenLib.py:181-204 model) -> (bases, lengths)."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    starts = rng.integers(0, G - L + 1, size=n)
    seg = genome[starts[:, None] + np.arange(L)[None, :]]
    thr_d, thr_i = int(p_del * 65536), int(p_ins * 65536)
    both = np.empty((n, 2 * L), dtype=np.uint8)
    both[:, 0::2] = seg
    both[:, 1::2] = _ACGT[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    mask = np.empty((n, 2 * L), dtype=bool)
    mask[:, 0::2] = rng.integers(0, 65536, size=(n, L), dtype=np.uint16) >= thr_d
    mask[:, 1::2] = rng.integers(0, 65536, size=(n, L), dtype=np.uint16) < thr_i
    return both[mask], mask.sum(axis=1)


def metagenome(n_genomes=10, G=5_000_000, seed=0):
    """Genomes (seed 1000 + g) and the chimeric contig pair of every adjacent genome pair (abunGap_example rule,
    dataGenerator.py:257-267: contigs = [g0[:G/2] + g1[G/2:], g1[:G/2] + g0[G/2:]])."""
    genomes = [random_genome(G, np.random.default_rng(1000 + seed * 1000 + g)) for g in range(n_genomes)]
    half = G // 2
    contigs = []
    for g in range(0, n_genomes - 1, 2):
        a, b = genomes[g], genomes[g + 1]
        contigs += [np.concatenate([a[:half], b[half:]]), np.concatenate([b[:half], a[half:]])]
    if n_genomes % 2:
        contigs.append(genomes[-1])
    return genomes, contigs


def lognormal_abundances(n_genomes, G, total_bases, lo=5.0, hi=200.0, seed=3000):
    """Coverage per genome: clip(lognormal, lo, hi) rescaled so that sum(c_i * G) ~ total_bases (config 3)."""
    c = np.clip(np.random.default_rng(seed).lognormal(mean=3.0, sigma=1.0, size=n_genomes), lo, hi)
    return c * (total_bases / (c.sum() * G))


def metagenome_read_counts(n_genomes, G, L, total_bases, lognormal=True, seed=3000):
    if lognormal:
        cov = lognormal_abundances(n_genomes, G, total_bases, seed=seed)
    else:
        cov = np.full(n_genomes, total_bases / (n_genomes * G))
    return np.maximum(1, (cov * G / L).astype(np.int64))


def metagenome_reads(genomes, counts, L, p, first, last, seed=0, threads=None):
    """Reads [first, last) of the global order (genome 0's counts[0] reads, then genome 1's, ...) as (bases, offsets).
    Chunks are generated on a thread pool (numpy releases the GIL in the heavy calls)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    jobs = []
    g_first = 0
    for g, n_g in enumerate(counts):
        lo, hi = max(first, g_first), min(last, g_first + int(n_g))
        if lo < hi:
            for c in range((lo - g_first) // CHUNK, (hi - g_first + CHUNK - 1) // CHUNK):
                n = min(CHUNK, int(n_g) - c * CHUNK)
                jobs.append((g, c, n, max(lo - g_first - c * CHUNK, 0), min(hi - g_first - c * CHUNK, n)))
        g_first += int(n_g)

    def one(job):
        g, c, n, a, z = job
        b, ln = _reads_chunk(genomes[g], n, L, p, p, (2000 + seed * 1000 + g) * 100003 + c)
        o = np.concatenate([[0], np.cumsum(ln)])
        return b[o[a]:o[z]], ln[a:z]
    with ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, jobs))
    lens = np.concatenate([r[1] for r in res]) if res else np.zeros(0, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return (np.concatenate([r[0] for r in res]) if res else np.zeros(0, dtype=np.uint8)), off


_RC = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _RC[_a] = _b


def doubled(bases, off):
    """IORobot.writeToFile_Double1 on arrays (IORobot.py:107-140): record 2i = read i, record 2i+1 = its reverse complement."""
    n = len(off) - 1
    lens = np.diff(off)
    off2 = np.zeros(2 * n + 1, dtype=np.int64)
    np.cumsum(np.repeat(lens, 2), out=off2[1:])
    out = np.empty(int(off2[-1]), dtype=np.uint8)
    for i in range(n):
        s = bases[off[i]:off[i + 1]]
        out[off2[2 * i]:off2[2 * i + 1]] = s
        out[off2[2 * i + 1]:off2[2 * i + 2]] = _RC[s[::-1]]
    return out, off2
