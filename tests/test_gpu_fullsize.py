"""Parity at BASELINE size: the references of configs 1 and 2 at their stated size (2 x 5 Mbp chimeric contigs; one 5 Mbp
contig), 3000 reads each, GPU rows against the CPU oracle run on all host threads (reference indexed once and shared).
bench.py does the same on its own batch and reports it as `parity`."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from metafinishersc_b200 import datagen, engine

pytestmark = pytest.mark.gpu


def _oracle_rows(O, contigs, reads, threads):
    ref = O.Ref(contigs)
    n = len(reads)
    bounds = [(n * t // threads, n * (t + 1) // threads) for t in range(threads)]
    bounds = [(a, b) for a, b in bounds if b > a]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda ab: O.align(ref, reads[ab[0]:ab[1]], band_cap=248), bounds))
    rows, cells, cap = [], 0, 0
    for (a, _), o in zip(bounds, res):
        r = o["rows"][:, :9].copy()
        r[:, 1] += a
        rows.append(r)
        cells += o["stats"]["cells"]
        cap += o["stats"]["cap_hits"]
    ref.free()
    return np.concatenate(rows), cells, cap


def _gpu_rows(E, contigs, reads):
    P = engine.make_params()
    ix = E.index(contigs, P)
    r = E.align(ix, reads, P)
    ix.free()
    got = np.stack([r.ref_id, r.qry_id, (r.s2 > r.e2).astype(np.int32), r.s1, r.e1, r.s2, r.e2, r.errors, r.cols], axis=1).astype(np.int64)
    return got, r.stats


def test_cfg1_full_reference_3000_reads(gpu_engine, oracle):
    d = datagen.abun_gap(G=5_000_000, repeat_len=12000, L=6000, p=0.01, abun=(0.9, 0.9), seed=0)      # 2 x 750 reads ...
    d2 = datagen.abun_gap(G=5_000_000, repeat_len=12000, L=6000, p=0.01, abun=(0.9, 0.9), seed=0)
    reads = d["reads"] + [r.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1] for r in d2["reads"]]     # ... and their reverse complements
    assert len(reads) == 3000
    got, st = _gpu_rows(gpu_engine, d["contigs"], reads)
    want, cells, cap = _oracle_rows(oracle, d["contigs"], reads, os.cpu_count() or 1)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert st["cells"] == cells and st["cap_hits"] == cap == 0
    assert len(got) >= 3000


def test_cfg2_full_reference_3000_reads(gpu_engine, oracle):
    d = datagen.sc_lr(G=5_000_000, n_reads=3000, seed=100, read_seed=2100)
    got, st = _gpu_rows(gpu_engine, d["contigs"], d["reads"])
    want, cells, cap = _oracle_rows(oracle, d["contigs"], d["reads"], os.cpu_count() or 1)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert st["cells"] == cells and st["cap_hits"] == cap == 0
    assert len(got) >= 2900
