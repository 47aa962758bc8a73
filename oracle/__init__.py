"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/mfsc_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  PARITY UNPINNED against real MUMmer 3.23 (absent from
/root/reference and from this machine); pinned only against the structure-determined
show-coords tables of the reference README.md:111-129.
"""
