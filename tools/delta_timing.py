#!/usr/bin/env python
"""Stage-4 time with and without the delta (indel list) recording, for both DP engines (cfg2 reads)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metafinishersc_b200 import datagen, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
d = datagen.sc_lr(G=5_000_000, n_reads=n, seed=100, read_seed=2100)
E = engine.Engine()
for eng in (0, 1):
    P = engine.make_params(engine=eng)
    ix = E.index(d["contigs"], P)
    rd = E.upload(engine._as_buf(d["reads"]))
    for flags, name in ((0, "rows"), (engine.WANT_DELTAS, "rows+deltas")):
        for _ in range(2):
            r = E.align_reads(ix, rd, P, flags=flags)
        print("engine %d %-12s extend %.1f ms  rows %d  handed over %d" % (eng, name, r.ms["extend"], len(r), r.stats["dp_retry_reads"]))
