#!/usr/bin/env python
"""A longer run of the seeded random parity scenarios of tests/test_emu_random.py on the GPU (seeds 60..459, or `first last` from the
command line): library through the C ABI against the oracle, rows-only and delta mode.  The test suite runs seeds 0..59."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases, test_emu_random
from metafinishersc_b200 import engine
from oracle import oracle as O
O.build()
E = engine.Engine()
bad_total = 0
first, last = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (60, 460)
for seed in range(first, last):
    contigs, reads, opts = test_emu_random.scenario(seed)
    bad, r, o = cases.run_case(E, O, contigs, reads, opts)
    if bad:
        bad_total += 1
        print("seed", seed, opts, bad[:2])
print("sweep done, seeds %d..%d, failures:" % (first, last - 1), bad_total)
