#!/usr/bin/env python
"""Where the time of alignerRobot.largeRvsQAlign goes on the cfg2 batch (cProfile, host clock).  usage: tools/wrapper_profile.py [0|1 fullDelta]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metafinishersc_b200 import alignerRobot, datagen, fasta, houseKeeper  # noqa: E402

alignerRobot.fullDelta = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
d = datagen.sc_lr(G=5_000_000, n_reads=25000, seed=100, read_seed=2100)
with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
    folder = tmp + "/"
    fasta.write_fasta(folder + "LC.fasta", ["Contig0"], d["contigs"])
    fasta.write_fasta(folder + "LR.fasta", datagen.names("Segkk", len(d["reads"])), d["reads"])
    alignerRobot.largeRvsQAlign(folder, 20, "", "LC", "LR", "warm")
    t0 = time.perf_counter()
    alignerRobot.largeRvsQAlign(folder, 20, "", "LC", "LR", "plain")
    print("plain run: %.3f s" % (time.perf_counter() - t0))
    pr = cProfile.Profile()
    pr.enable()
    alignerRobot.largeRvsQAlign(folder, 20, "", "LC", "LR", "prof")
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
