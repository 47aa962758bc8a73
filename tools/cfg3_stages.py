#!/usr/bin/env python
"""Stage split on a BASELINE config 3 / 5 shaped case: many 5 Mbp genomes as contigs, 10 kbp reads at 2 % indel error drawn
from them with log-normal abundances.  usage: cfg3_stages.py [n_genomes] [n_reads] [kmer]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metafinishersc_b200 import datagen, engine  # noqa: E402

n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
G = 5_000_000
genomes = [datagen.random_genome(G, np.random.default_rng(1000 + g)) for g in range(n_gen)]
ab = np.clip(np.random.default_rng(3000).lognormal(3.0, 1.0, size=n_gen), 5, 200)
counts = np.maximum(1, (ab / ab.sum() * n_reads).astype(int))
reads = []
for g in range(n_gen):
    r, _ = datagen.noisy_reads(genomes[g], int(counts[g]), 10_000, 0.01, 0.01, np.random.default_rng(2000 + g))
    reads += r
E = engine.Engine()
P = engine.make_params(kmer=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
ix = E.index([x.tobytes() for x in genomes], P)
rd = E.upload(engine._as_buf(reads))
for _ in range(3):
    r = E.align_reads(ix, rd, P)
bases = sum(map(len, reads))
print("%d genomes (%d Mbp, k=%d, index %.0f ms), %d reads (%.0f Mbp): %d rows, %.0f Mbp/s" %
      (n_gen, n_gen * G // 1_000_000, ix.info()["k"], ix.build_ms(), len(reads), bases / 1e6, len(r), bases / r.ms["total"] / 1e3))
print("stage ms:", {k: round(float(v), 2) for k, v in r.ms.items()}, "cells %.3g" % r.stats["cells"])
