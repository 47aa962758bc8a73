import sys, os
sys.path.insert(0, os.getcwd())
from metafinishersc_b200 import datagen, engine
E = engine.Engine()
d = datagen.sc_lr(G=5_000_000, n_reads=1250, seed=100, read_seed=2100)
P = engine.make_params()
ix = E.index(d["contigs"], P)
for _ in range(3):
    r = E.align(ix, d["reads"], P)
print({k: r.stats[k] for k in ("cells", "dp_calls", "gap_tasks", "gap_hits", "max_band", "dp_retry_reads")}, len(r), {k: round(v, 2) for k, v in r.ms.items()})
