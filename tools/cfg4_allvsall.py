#!/usr/bin/env python
"""BASELINE config 4 shape at reduced size: reads vs reads, both written through writeToFile_Double1 (IORobot.py:107-140:
every read followed by its reverse complement), nucmer --maxmatch.  Prints stage times and checks size-independent
properties: every record aligns to itself end to end, and the row set is symmetric under swapping reference and query."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metafinishersc_b200 import datagen, engine  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
comp = bytes.maketrans(b"ACGT", b"TGCA")
g = datagen.random_genome(G, np.random.default_rng(1000))
reads, _ = datagen.noisy_reads(g, n_reads, 10_000, 0.01, 0.01, np.random.default_rng(2000))
dbl = []
for r in reads:
    dbl += [r, r.translate(comp)[::-1]]
E = engine.Engine()
P = engine.make_params()
t0 = time.perf_counter()
ix = E.index(dbl, P)
t1 = time.perf_counter()
E.align(ix, dbl, P)                # warm-up: pool growth, first touch
t1b = time.perf_counter()
r = E.align(ix, dbl, P)
t2 = time.perf_counter()
align_s = t2 - t1b
print("records %d (%.1f Mbp), index %.2f s (k=%d), align %.2f s: %d rows, %d matches, %.3g DP cells, handed over %d"
      % (len(dbl), sum(map(len, dbl)) / 1e6, t1 - t0, ix.info()["k"], align_s, len(r), r.stats["mems"], r.stats["cells"], r.stats["dp_retry_reads"]))
print("stage ms:", {k: round(v, 1) for k, v in r.ms.items()})
lens = np.array([len(x) for x in dbl])
fwd = r.s2 < r.e2
selfrows = (r.ref_id == r.qry_id) & fwd & (r.s1 == 1) & (r.e1 == lens[r.ref_id]) & (r.s2 == 1) & (r.e2 == lens[r.qry_id]) & (r.errors == 0)
assert len(np.unique(r.ref_id[selfrows])) == len(dbl), "some record does not align to itself end to end"
# symmetry: (a, b, s1, e1, s2, e2) forward rows <-> (b, a, s2, e2, s1, e1)
A = set(zip(r.ref_id[fwd].tolist(), r.qry_id[fwd].tolist(), r.s1[fwd].tolist(), r.e1[fwd].tolist(), r.s2[fwd].tolist(), r.e2[fwd].tolist()))
B = set((b, a, s2, e2, s1, e1) for (a, b, s1, e1, s2, e2) in A)
print("forward rows %d, symmetric counterpart present for %.2f %%" % (len(A), 100.0 * len(A & B) / max(1, len(A))))
