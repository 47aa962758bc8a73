#!/usr/bin/env python
"""Both stage-4 DP engines on reads with substitution errors (which the reference's generator never makes): rows must be
identical, and the packed engine should not have to hand reads over.  Prints cells, stage time and hand-over count."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metafinishersc_b200 import datagen, engine
E=engine.Engine()
rng=np.random.default_rng(7)
for psub, pid in ((0.02,0.05),(0.05,0.05),(0.10,0.02),(0.15,0.0)):
    d=datagen.sc_lr(G=2_000_000, n_reads=3000, seed=100, read_seed=2100, p_del=pid, p_ins=pid)
    reads=[]
    for r in d["reads"]:
        a=np.frombuffer(r,dtype=np.uint8).copy()
        m=rng.random(len(a))<psub
        a[m]=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,size=int(m.sum()))]
        reads.append(a.tobytes())
    out=[]
    for eng in (0,1):
        P=engine.make_params(engine=eng)
        ix=E.index(d["contigs"],P)
        r=E.align(ix,reads,P)
        out.append(r)
        print("psub %.2f pindel %.2f engine %d rows %d cells %d extend %.1f ms handed over %d"%(psub,2*pid,eng,len(r),r.stats["cells"],r.ms["extend"],r.stats["dp_retry_reads"]))
    same=all(np.array_equal(getattr(out[0],f),getattr(out[1],f)) for f in ("ref_id","qry_id","s1","e1","s2","e2","errors","cols"))
    print("   engines agree:", same)
