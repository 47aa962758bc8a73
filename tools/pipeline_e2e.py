#!/usr/bin/env python
"""File-level timing of the drop-in wrapper: FASTA on disk -> largeRvsQAlign -> rows (what mFixer calls), cfg2 shape."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metafinishersc_b200 import alignerRobot, datagen, fasta, coverage, IORobot

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
full = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
d = datagen.sc_lr(G=5_000_000, n_reads=n_reads, seed=100, read_seed=2100)
folder = tempfile.mkdtemp(dir="/dev/shm") + "/"
fasta.write_fasta(folder + "LC_filtered.fasta", ["Segkk0"], d["contigs"])
fasta.write_fasta(folder + "SR.fasta", datagen.names("Segkk", len(d["reads"])), d["reads"])
alignerRobot.fullDelta = full
alignerRobot.get_engine()
for rep in range(2):
    alignerRobot.clear_caches()
    t0 = time.time()
    rows = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", "SR2LC_filteredAlign")
    t1 = time.time()
    print("rep %d: largeRvsQAlign %.2f s, %d rows, %.1f Mbp/s (fullDelta=%s)" % (rep, t1 - t0, len(rows), sum(len(r) for r in d["reads"]) / (t1 - t0) / 1e6, full))
t0 = time.time()
cov = coverage.alignSR2LC(folder, "", "LC_filtered")
print("alignSR2LC (align + select + coverage + json) %.2f s" % (time.time() - t0))
shutil.rmtree(folder)
