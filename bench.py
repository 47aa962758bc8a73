#!/usr/bin/env python
"""bench.py -- read->contig alignment throughput (aligned Mbp/s) of the hot path on N B200s.

A step = one pass of stages 2-4 (seeds, clusters, extension) over one batch of synthetic reads against a contig index
resident in HBM.

Primary line (`value`, every N): BASELINE.json configs[1] -- one 5 Mbp genome, 50X of 10 kbp reads with 15 % indel error
(25 000 reads, ~250 Mbp) PER GPU: reads shard by batch with no data-path collective, so the contract's `"scaling": "weak"`.
  value       : Mbp/s with the read batch already packed in HBM (CUDA events inside the library around the step)
  e2e         : the same metric through Engine.align (mfsc_align_batch; pieces via mfsc_reads_upload + mfsc_align_reads for a large batch) from pinned HOST buffers, timed with the HOST clock around the
                call a user makes (H2D of the reads, D2H of the rows, ctypes marshalling all inside)
  e2e_wrapper : the same reads as FASTA files through the reference-facing alignerRobot.largeRvsQAlign (split into 20
                parts, 20 nucmer jobs, .delta and coords files written, rows returned), N = 1 only
  strong      : BASELINE.json configs[2] as ONE fixed job sharded over the N ranks -- 10 x 5 Mbp chimeric contigs,
                ~150 000 reads x 10 kbp (1.5 Gbp) at 1 % error dealt to the ranks in contiguous batches; index build on rank 0
                + NCCL broadcast (mfsc_index_broadcast) + read upload + alignment + rows gathered on rank 0 in rank order,
                ALL inside the measured wall (host clock, barrier on both sides, max over ranks)
  roofline*   : every stage against its bound; `traffic` = ncu dram bytes per launch of the same command
                (profiles/r02_traffic.json)
  cpu_baseline, parity : the CPU oracle on all host threads over the first reads of the batch, and its rows compared
                with the GPU rows of the same reads (the run fails when they differ)
`--workload cfg1|cfg3|cfg4|cfg5` selects another BASELINE shape for the primary line; `--scaling strong` makes the
sharded job the primary line.  `--impl reference`: the CPU oracle (restatement of nucmer+show-coords; MUMmer itself is
not in the image, see "mummer" in the line) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from metafinishersc_b200 import datagen  # noqa: E402

METRIC = "read->contig aligned Mbp/s (nucmer --maxmatch path: seeds+clusters+extension)"
OPS_PER_CELL = 12      # SASS of the packed cell (k_extend<false,true>): 1 VIMNMX3 + 2 VIMNMX + 3 LOP3 + 5 IMAD + 1 predicated IMAD (profiles/r02_extend_sass.md)

WORKLOADS = {
    "cfg1": "cfg1: README shape, 2 x 5 Mbp chimeric contigs, 58332 reads x 6 kbp, 1% del + 1% ins, per GPU",
    "cfg2": "cfg2: 5 Mbp genome, 50X, 10 kbp reads, 15% indel (p_del=p_ins=0.075), 25000 reads/GPU",
    "cfg3": "cfg3: 10 genomes x 5 Mbp (10 chimeric contigs), log-normal 5-200X, %d reads x 10 kbp, 1%% del + 1%% ins, per GPU",
    "cfg4": "cfg4: first %d reads (10 kbp, 1%% error) of the cfg5 read set, doubled by writeToFile_Double1 (Read_p + Read_d), aligned against themselves, per GPU",
    "cfg5": "cfg5: 100 genomes x 5 Mbp (500 Mbp of contigs), uniform 30X; a %d-read sample (x 10 kbp, 1%% error) of the 1.5 M reads per step, per GPU",
}


def make_workload(rank, args):
    """-> (contigs, reads, label).  contigs / reads: list[bytes] or (uint8 bases, int64 offsets)."""
    w = args.workload
    if w == "cfg1":      # README shape: two 5 Mbp genomes at 50X/20X, 6 kbp reads at 1% del + 1% ins, 2 chimeric contigs
        d = datagen.abun_gap(G=args.genome, repeat_len=12000, L=6000, p=0.01, abun=(50, 20), seed=rank)
        g = datagen.abun_gap(G=args.genome, repeat_len=12000, L=6000, p=0.01, abun=(0.01, 0.01), seed=0) if rank else d
        reads = d["reads"][:args.reads] if args.reads and args.reads < len(d["reads"]) else d["reads"]
        return g["contigs"], reads, WORKLOADS[w]
    if w == "cfg2":
        n = args.reads or 25000
        d = datagen.sc_lr(G=args.genome, n_reads=n, seed=100, read_seed=2100 + rank)
        label = WORKLOADS[w] if (n == 25000 and args.genome == 5_000_000) else "cfg2 shape scaled: %d bp genome, %d reads/GPU" % (args.genome, n)
        return d["contigs"], d["reads"], label
    if w == "cfg3":
        genomes, contigs = datagen.metagenome(10, args.genome)
        counts = datagen.metagenome_read_counts(10, args.genome, 10000, 1.5e9 * args.genome / 5_000_000)
        n = int(counts.sum()) if not args.reads else min(args.reads, int(counts.sum()))
        reads = datagen.metagenome_reads(genomes, counts, 10000, 0.01, 0, n, seed=rank)
        return [c.tobytes() for c in contigs], reads, WORKLOADS[w] % n
    if w == "cfg5":
        genomes, contigs = datagen.metagenome(100, args.genome)
        counts = datagen.metagenome_read_counts(100, args.genome, 10000, 30.0 * 100 * args.genome, lognormal=False)
        n = args.reads or 30000
        # a sample spread over all genomes: every (total / n)-th chunk of the global read order
        per = max(1, n // 100)
        parts = [datagen.metagenome_reads(genomes, counts, 10000, 0.01, int(counts[:g].sum()), int(counts[:g].sum()) + per, seed=rank) for g in range(100)]
        buf = np.concatenate([p[0] for p in parts])
        off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(p[1]) for p in parts]))]).astype(np.int64)
        return [c.tobytes() for c in contigs], (buf, off), WORKLOADS[w] % (len(off) - 1)
    if w == "cfg4":
        genomes, _ = datagen.metagenome(100, args.genome)
        counts = datagen.metagenome_read_counts(100, args.genome, 10000, 30.0 * 100 * args.genome, lognormal=False)
        n = args.reads or 20000
        # "first n reads" of a shuffled cfg5 read set = n reads spread uniformly over the genomes
        per = max(1, n // 100)
        parts = [datagen.metagenome_reads(genomes, counts, 10000, 0.01, int(counts[:g].sum()), int(counts[:g].sum()) + per, seed=rank) for g in range(100)]
        buf = np.concatenate([p[0] for p in parts])
        off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(p[1]) for p in parts]))]).astype(np.int64)
        dbl = datagen.doubled(buf, off)
        return dbl, dbl, WORKLOADS[w] % (len(off) - 1)
    raise SystemExit("unknown workload " + w)


def as_list(seqs):
    if isinstance(seqs, tuple):
        buf, off = seqs
        return [buf[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
    return seqs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.lines, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.p:
            self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_copy(E, arr):
    """numpy array over pinned host memory holding a copy of arr"""
    p = ctypes.c_void_p()
    nbytes = max(int(arr.nbytes), 16)
    rc = E.L.mfsc_host_alloc(ctypes.byref(p), nbytes)
    if rc != 0:
        return arr
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size)
    out[:] = arr.ravel()
    return out


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host threads, reference indexed once and shared
# ------------------------------------------------------------------------------------------
def oracle_run(ref, reads, threads, minmatch=20):
    """Aligns `reads` (list of bytes) in contiguous per-thread chunks.  -> (rows int64 [n, 9] in read order, seconds, cells, stats)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    n = len(reads)
    bounds = [(n * t // threads, n * (t + 1) // threads) for t in range(threads)]
    bounds = [(a, b) for a, b in bounds if b > a]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda ab: O.align(ref, reads[ab[0]:ab[1]], band_cap=248, minmatch=minmatch), bounds))
    dt = time.perf_counter() - t0
    rows = []
    for (a, _), o in zip(bounds, res):
        r = o["rows"][:, :9].copy()
        r[:, 1] += a
        rows.append(r)
    rows = np.concatenate(rows) if rows else np.zeros((0, 9), dtype=np.int64)
    st = {"cells": sum(o["stats"]["cells"] for o in res), "max_band": max([o["stats"]["max_band"] for o in res] or [0]),
          "cap_hits": sum(o["stats"]["cap_hits"] for o in res)}
    return rows, dt, st


def size_sample(ref, reads, threads, target_s):
    """Reads per sample so that the oracle runs about target_s on `threads` threads (two pilots, linear fit)."""
    n1, n2 = 2 * threads, 8 * threads
    _, t1, _ = oracle_run(ref, reads[:n1], threads)
    _, t2, _ = oracle_run(ref, reads[:n2], threads)
    b = max((t2 - t1) / (n2 - n1), 1e-6)
    a = max(t1 - b * n1, 0.0)
    return int(max(threads, min(len(reads), (target_s - a) / b)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import mummer_probe
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    contigs, reads, label = make_workload(0, args)
    contigs, reads = as_list(contigs), as_list(reads)
    ref = O.Ref(contigs)                      # indexed once (outside the timed steps, like the GPU arm's resident index)
    step_s = min(args.ref_step_seconds, 150.0 / max(1, args.steps + args.warmup))
    n_sample = size_sample(ref, reads, threads, step_s)
    reads = reads[:n_sample]
    for _ in range(args.warmup):
        oracle_run(ref, reads[:threads], threads)
    total_dt, total_bases, cells = 0.0, 0, 0
    for _ in range(args.steps):
        _, dt, st = oracle_run(ref, reads, threads)
        total_dt += dt; total_bases += sum(len(x) for x in reads); cells += st["cells"]
    v = total_bases / total_dt / 1e6
    sample = "%d reads (%.1f Mbp) per step of the batch, CPU oracle (restatement of nucmer --maxmatch + show-coords, not MUMmer; " \
             "reference indexed once and shared by the threads)" % (n_sample, total_bases / args.steps / 1e6)
    emit(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": total_dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                      "config": {"workload": label, "sample": sample},
                      "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": threads, "kind": "port", "sample": sample,
                                       "dp_gcups": cells / total_dt / 1e9},
                      "mummer": mummer_probe.report(),
                      "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (NCCL's
    version banner, library chatter) has been diverted to stderr by main()."""
    os.write(_STDOUT_FD, (line + "\n").encode())


_STDOUT_FD = 1


def rows_matrix(r):
    if len(r) == 0:
        return np.zeros((0, 9), dtype=np.int64)
    return np.stack([r.ref_id, r.qry_id, (r.s2 > r.e2).astype(np.int32), r.s1, r.e1, r.s2, r.e2, r.errors, r.cols], axis=1).astype(np.int64)


# ------------------------------------------------------------------------------------------
# the sharded fixed job (strong scaling, BASELINE configs[2])
# ------------------------------------------------------------------------------------------
def measure_strong(E, args, rank, world, local, comm, torch, dist):
    from metafinishersc_b200 import engine
    G = args.strong_genome
    genomes, contigs = datagen.metagenome(10, G)
    counts = datagen.metagenome_read_counts(10, G, 10000, 1.5e9 * G / 5_000_000)
    n_total = int(counts.sum())
    a, b = n_total * rank // world, n_total * (rank + 1) // world      # contiguous batches in the global read order
    buf, off = datagen.metagenome_reads(genomes, counts, 10000, 0.01, a, b)
    cbuf, coff = engine._as_buf([c.tobytes() for c in contigs])
    hbuf, hoff, hcb, hco = pinned_copy(E, buf), pinned_copy(E, off), pinned_copy(E, cbuf), pinned_copy(E, coff)
    P = engine.make_params()
    bases_total = None
    t_bases = torch.tensor([float(off[-1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_bases)
    bases_total = float(t_bases.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gbuf = {}
    if world > 1:
        cap = 2 * (n_total // world + 1)
        gbuf = {"cap": cap, "cnt": torch.zeros(1, dtype=torch.int64, device="cuda"), "all": torch.zeros(world, dtype=torch.int64, device="cuda"),
                "mine": torch.zeros((cap, 8), dtype=torch.int32, device="cuda"), "stage": torch.zeros((4 * cap, 8), dtype=torch.int32).pin_memory(),
                "list": [torch.zeros((cap, 8), dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None}

    n_local = len(off) - 1
    # pieces of about 100 Mbp (a piece costs about a millisecond of fixed work: launches, syncs, sort set-up), at most --strong-pieces
    n_pieces = max(1, min(args.strong_pieces, n_local, int(round(float(off[-1]) / 100e6))))
    if args.strong_pieces_exact:
        n_pieces = max(1, min(args.strong_pieces_exact, n_local))
    cuts = [n_local * k // n_pieces for k in range(n_pieces + 1)]

    def one_step():
        import queue
        parts = {}
        box = {}
        q = queue.Queue()

        def up():
            # the rank's read batch goes up (and is packed) in pieces on a second host thread / stream, while the index is made
            # and while the pieces before it are aligned
            try:
                E2 = engine.Engine(E.L)
                E2.set_device(local)
                for k in range(n_pieces):
                    lo, hi = cuts[k], cuts[k + 1]
                    q.put(E2.upload((hbuf[hoff[lo]:hoff[hi]], hoff[lo:hi + 1] - hoff[lo])))
                box["upload_done"] = time.perf_counter()
                E2.thread_release()
            except Exception as e:       # surfaced below
                box["err"] = e
                q.put(None)
        barrier()
        t0 = time.perf_counter()
        th = threading.Thread(target=up)
        th.start()
        ix = None
        if world > 1 and args.strong_index == "sharded":
            # every rank builds 1/N of the k-mer space, the slices are exchanged over NVLink (mfsc_index_build_sharded)
            ix, bms = comm.build_sharded((hcb, hco), P)
            parts["index_sharded_upload_ms"], parts["index_sharded_slice_ms"], parts["index_sharded_exchange_ms"] = bms[0], bms[1], bms[2]
            parts["index_build_ms"] = (time.perf_counter() - t0) * 1e3
            parts["broadcast_ms"] = 0.0
        else:
            if rank == 0:
                ix = E.index((hcb, hco), P)
                parts["index_build_ms"] = (time.perf_counter() - t0) * 1e3
            tb = time.perf_counter()
            if world > 1:
                ix, _ = comm.broadcast(ix, len(hco) - 1, root=0)
            parts["broadcast_ms"] = (time.perf_counter() - tb) * 1e3 if world > 1 else 0.0
        ta = time.perf_counter()
        rows_l, stage_ms, wait_ms = [], {}, 0.0
        for k in range(n_pieces):
            tw = time.perf_counter()
            reads_k = q.get()
            wait_ms += (time.perf_counter() - tw) * 1e3
            if reads_k is None:
                raise box["err"]
            rk = E.align_reads(ix, reads_k, P)
            reads_k.free()
            rk.qry_id = rk.qry_id + cuts[k]
            rows_l.append(rk)
            for kk, v in rk.ms.items():
                stage_ms[kk] = stage_ms.get(kk, 0.0) + v
        th.join()
        parts["align_ms"] = (time.perf_counter() - ta) * 1e3
        parts["waited_for_uploads_ms"] = wait_ms
        parts["upload_done_ms"] = (box["upload_done"] - t0) * 1e3

        class _R:
            pass
        r = _R()
        for fld in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
            setattr(r, fld, np.concatenate([getattr(x, fld) for x in rows_l]))
        r.ms = stage_ms
        n_rows_local = len(r.s1)
        tg = time.perf_counter()
        n_rows = n_rows_local
        if world > 1:
            # rows to rank 0 in rank order (= the global read order): the counts with one all-gather, then one gather of
            # fixed-size device buffers (capacity: 2 rows per read of the largest shard, grown when a step needs more)
            cap = gbuf["cap"]
            gbuf["cnt"][0] = n_rows
            dist.all_gather_into_tensor(gbuf["all"], gbuf["cnt"])
            counts_h = gbuf["all"].cpu().tolist()
            if max(counts_h) > cap:
                cap = gbuf["cap"] = 2 * max(counts_h)
                gbuf["mine"] = torch.zeros((cap, 8), dtype=torch.int32, device="cuda")
                gbuf["stage"] = torch.zeros((cap, 8), dtype=torch.int32).pin_memory()
                gbuf["list"] = [torch.zeros_like(gbuf["mine"]) for _ in range(world)] if rank == 0 else None
            if n_rows:
                gbuf["stage"][:n_rows] = torch.from_numpy(np.stack([r.ref_id, r.qry_id + a, r.s1, r.e1, r.s2, r.e2, r.errors, r.cols], axis=1))
                gbuf["mine"][:n_rows].copy_(gbuf["stage"][:n_rows], non_blocking=True)
            dist.gather(gbuf["mine"], gbuf["list"], dst=0)
            if rank == 0:
                allrows = torch.cat([g[:c] for g, c in zip(gbuf["list"], counts_h)]).cpu().numpy()
                n_rows = len(allrows)
        parts["gather_ms"] = (time.perf_counter() - tg) * 1e3
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        parts["stage_ms"] = dict(r.ms)
        parts["rows_total"] = n_rows
        ix.free()
        return wall, parts

    for _ in range(max(1, args.warmup)):          # the scratch pool was trimmed before this phase: it grows back over the first steps
        one_step()
    K = max(1, min(args.steps, args.strong_steps))
    walls, last = [], None
    for _ in range(K):
        w, last = one_step()
        walls.append(w)
    wall = sum(walls)
    t = torch.tensor([wall, last["align_ms"], -last["align_ms"]] + walls, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, align_max, align_min = float(t[0].item()), float(t[1].item()), -float(t[2].item())
    walls_max = [float(x) for x in t[3:].tolist()]
    return {"workload": "cfg3 as one job: 10 chimeric contigs x %d bp, %d reads x 10 kbp (%.0f Mbp), 1%% del + 1%% ins, contiguous read "
                        "batches over %d rank(s)" % (G, n_total, bases_total / 1e6, world),
            "scaling": "strong", "value": bases_total * K / (wall_max * 1e-3) / 1e6, "unit": "Mbp/s", "steps": K, "ms_per_step": wall_max / K,
            "step_walls_ms_max_over_ranks": [round(x, 2) for x in walls_max], "ms_per_step_median": statistics.median(walls_max),
            "index": ("sharded build: every rank packs the contigs and builds 1/N of the k-mer space, slices exchanged with grouped NCCL broadcasts"
                      if world > 1 and args.strong_index == "sharded" else "built on rank 0, NCCL broadcast of the 5 index buffers") if world > 1 else "built in place",
            "pieces_per_rank": n_pieces,
            "inside_the_wall": "index (see `index`; from pinned host contigs) + H2D and packing of the rank's reads in `pieces_per_rank` pieces "
                               "(second host thread/stream, overlapped with the index build and with the alignment of the pieces before) + stages 2-4 + rows to the "
                               "host + rows gathered on rank 0 in rank order; host clock, barrier both sides, max over ranks",
            "rank0_breakdown_ms": {k: v for k, v in last.items() if k.endswith("_ms")}, "rank0_stage_ms": last["stage_ms"],
            "align_ms_max_over_ranks": align_max, "align_ms_min_over_ranks": align_min, "rows_total": last["rows_total"]}


# ------------------------------------------------------------------------------------------
# the reference-facing wrapper on files (N = 1)
# ------------------------------------------------------------------------------------------
def measure_wrapper(contigs, reads, n_bases):
    from metafinishersc_b200 import alignerRobot, fasta, houseKeeper
    _DEFAULT_FULL_DELTA = bool(alignerRobot.fullDelta)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as tmp:
        folder = tmp + "/"
        contigs, reads = as_list(contigs), as_list(reads)
        fasta.write_fasta(folder + "LC.fasta", datagen.names("Contig", len(contigs)), contigs)
        fasta.write_fasta(folder + "LR.fasta", datagen.names("Segkk", len(reads)), reads)
        cwd = os.getcwd()
        res = {}
        try:
            for mode, full in (("full_delta", True), ("rows_only", False)):
                alignerRobot.fullDelta = full
                best, rows = None, 0
                for it in range(3):            # first pass warms the index cache and the page cache, like the second job of a run
                    t0 = time.perf_counter()
                    data = alignerRobot.largeRvsQAlign(folder, houseKeeper.globalParallelFileNum, "", "LC", "LR", "bench" + mode + str(it))
                    dt = time.perf_counter() - t0
                    best = dt if best is None or dt < best else best
                    rows = len(data)
                res[mode] = (best, rows)
        finally:
            os.chdir(cwd)
            alignerRobot.fullDelta = _DEFAULT_FULL_DELTA
            alignerRobot.clear_caches()
    best, rows = res["full_delta" if _DEFAULT_FULL_DELTA else "rows_only"]
    return {"value": n_bases / best / 1e6, "unit": "Mbp/s", "seconds": best, "rows": rows,
            "call": "alignerRobot.largeRvsQAlign(folder, 20, mummerLink, 'LC', 'LR', header): FASTA on tmpfs -> 20 parts -> 20 jobs -> "
                    ".delta + coords files written -> rows (host clock, best of 3)", "full_delta": _DEFAULT_FULL_DELTA,
            "with_full_delta_records": {"value": n_bases / res["full_delta"][0] / 1e6, "seconds": res["full_delta"][0]},
            "with_alignment_headers_only": {"value": n_bases / res["rows_only"][0] / 1e6, "seconds": res["rows_only"][0]}}


def main():
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0: the workload's own size; cfg2: 25000)")
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--ref-step-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", type=int, default=0, help="stage-4 DP engine: 0 packed (+ general for handed-over reads), 1 general only")
    ap.add_argument("--kmer", type=int, default=0, help="index k-mer length (0: chosen from the reference size)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS), help="cfg2 is the headline; the others are BASELINE's other shapes")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: the sharded cfg3 job becomes the primary line")
    ap.add_argument("--no-strong", action="store_true", help="skip the sharded cfg3 job")
    ap.add_argument("--strong-only", action="store_true", help="only the sharded cfg3 job (experiments)")
    ap.add_argument("--no-wrapper", action="store_true", help="skip the file-level alignerRobot.largeRvsQAlign measurement")
    ap.add_argument("--strong-steps", type=int, default=6)
    ap.add_argument("--strong-pieces-exact", type=int, default=0, help="force the number of pieces (experiments)")
    ap.add_argument("--strong-pieces", type=int, default=4, help="sub-batches a rank's reads go up in (upload of piece k+1 overlaps the alignment of piece k)")
    ap.add_argument("--strong-index", default="sharded", choices=["sharded", "broadcast"],
                    help="N > 1: every rank builds 1/N of the index and the slices are exchanged (sharded), or rank 0 builds and broadcasts")
    ap.add_argument("--strong-genome", type=int, default=5_000_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from metafinishersc_b200 import engine, multigpu
    torch.cuda.set_device(local)
    E = engine.Engine()
    E.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def exchange(buf, src):
            t = torch.from_numpy(buf).cuda()
            dist.broadcast(t, src=src)
            buf[:] = t.cpu().numpy()
        comm = multigpu.RankComm(E, rank, world, exchange)
    if args.strong_only:
        strong = measure_strong(E, args, rank, world, local, comm, torch, dist)
        if rank == 0:
            emit(json.dumps({"metric": METRIC, "value": strong["value"], "unit": "Mbp/s", "n_gpus": world, "steps": strong["steps"], "warmup": args.warmup,
                             "ms_per_step": strong["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                             "data": "synthetic", "config": {"workload": strong["workload"]}, "strong": strong}))
        if comm is not None:
            comm.free()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    P = engine.make_params(engine=args.engine, kmer=args.kmer)
    contigs, reads, label = make_workload(rank, args)
    buf, off = engine._as_buf(reads)
    cbuf, coff = engine._as_buf(contigs)
    n_bases = int(off[-1])
    n_reads = len(off) - 1

    # stage 1 on rank 0, replicated with one NCCL broadcast per buffer (mfsc_index_broadcast)
    t0 = time.perf_counter()
    ix, bcast_ms = multigpu.replicated_index(E, (cbuf, coff), P, rank, world, comm)
    index_wall_ms = (time.perf_counter() - t0) * 1e3
    info = ix.info()

    hbuf, hoff = pinned_copy(E, buf), pinned_copy(E, off)
    resident = E.upload((buf, off))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.fill_(1)
        torch.cuda.synchronize()
        return E.align_reads(ix, resident, P)

    def step_e2e():
        flush.fill_(1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = E.align(ix, (hbuf, hoff), P)
        return r, (time.perf_counter() - t) * 1e3

    for _ in range(args.warmup):
        r = step_resident()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    dev_ms, stage = 0.0, {}
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_resident()
        dev_ms += r.ms["total"]
        launches += r.stats["launches"]
        for k, v in r.ms.items():
            stage[k] = stage.get(k, 0.0) + v
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    cells, n_rows = r.stats["cells"], len(r)
    seed_bytes = r.stats["seed_bytes"]

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    e2e_ms, e2e_dev_ms = 0.0, 0.0
    for _ in range(args.steps):
        re, host_ms = step_e2e()
        e2e_ms += host_ms                     # host clock around the call a user makes
        e2e_dev_ms += re.ms["total"]
    barrier()
    clk = clocks.stop() if rank == 0 else None
    h2d, d2h = re.stats["h2d_bytes"], re.stats["d2h_bytes"]

    def reduce_(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_(x, dist.ReduceOp.MAX)

    def sum_over_ranks(x):
        return reduce_(x, dist.ReduceOp.SUM)

    dev_ms_max = max_over_ranks(dev_ms)
    e2e_ms_max = max_over_ranks(e2e_ms)
    wall_ms_max = max_over_ranks(wall_ms)
    bases_all = sum_over_ranks(float(n_bases))
    cells_all = sum_over_ranks(float(cells))
    ext_ms_max = max_over_ranks(stage["extend"])      # every collective happens on all ranks, before the rank-0 block
    max_band_all = max_over_ranks(float(r.stats["max_band"]))
    cap_hits_all = sum_over_ranks(float(r.stats["cap_hits"]))

    # stage 5 on the rows of the last step (rank 0's rows): coverage profile of the contigs, kept resident
    cov_ms, cov_bytes = None, None
    if rank == 0 and n_rows:
        sel = np.ones(n_rows, dtype=bool)
        prof = E.coverage_profile(r.ref_id[sel], r.s1[sel], r.e1[sel], coff)
        prof.free()
        prof = E.coverage_profile(r.ref_id[sel], r.s1[sel], r.e1[sel], coff)
        cov_ms = float(prof.ms[0])
        cov_bytes = 8 * n_rows + 12 * int(coff[-1])        # 8 N_rows + 4 C diff written + 4 C read + 4 C written
        prof.free()

    # the sharded fixed job
    resident.free()
    del flush
    torch.cuda.empty_cache()
    E.trim_pool()            # the sharded job starts from an empty scratch pool, as it would in a process of its own
    strong = None
    if not args.no_strong:
        strong = measure_strong(E, args, rank, world, local, comm, torch, dist)

    K = args.steps
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except Exception:
            pass
        tr = traffic.get(args.workload, {}) if not args.reads and args.genome == 5_000_000 and not args.kmer else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        ext_ms = stage["extend"] / K
        seed_ms = stage["seeds"] / K
        clu_ms = stage["cluster"] / K
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        alu_peak = n_sm * 128 * sm_mhz * 1e6 / 1e12   # Tiop/s at the clock observed in this run
        ext_tiops = cells * OPS_PER_CELL / (ext_ms * 1e-3) / 1e12 if ext_ms > 0 else 0.0
        issue_peak = n_sm * 4 * sm_mhz * 1e6          # warp instructions per second: 4 schedulers per SM, one per clock

        def traffic_of(kernel):
            t = tr.get(kernel) or {}
            return (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None

        def hbm_line(kernel, alg_bytes, ms, note):
            ach = alg_bytes / (ms * 1e-3) / 1e9 if ms and ms > 0 else None
            return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak if ach is not None else None, "algorithmic_bytes": alg_bytes, "traffic": traffic_of(kernel),
                    "peak_source": peak_src, "note": note}
        ext_tr = tr.get("k_extend") or {}
        M = r.stats["mems"]
        tables_ms = ix.build_ms(tables_only=True)
        R_bases = info["bases"]
        out = {
            "metric": METRIC, "value": bases_all * K / (dev_ms_max * 1e-3) / 1e6, "unit": "Mbp/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": label, "reads_per_gpu": n_reads, "query_bases_per_gpu": n_bases, "l2": "flushed (256 MiB write) between steps",
                       "dp_engine": "packed 32-bit states, general engine for handed-over reads" if args.engine == 0 else "general",
                       "index_k": info["k"], "index_hbm_bytes": info["hbm_bytes"], "parallelism": "reads sharded by batch x%d" % world},
            "wall_ms_per_step": wall_ms_max / K,
            "stage_ms_per_step": {k: v / K for k, v in stage.items()},
            "dp_gcups": cells_all / (ext_ms_max / K * 1e-3) / 1e9 if ext_ms_max > 0 else None,
            "dp_cells_per_step": cells_all, "rows_per_step_rank0": n_rows, "dp_reads_handed_over_rank0": r.stats.get("dp_retry_reads"),
            "max_band": int(max_band_all), "band_cap": 248, "diagonals_where_the_cap_bound": int(cap_hits_all),
            "index_build_ms": ix.build_ms() if world == 1 or rank == 0 else None, "index_tables_ms": tables_ms,
            "index_broadcast_ms": bcast_ms, "index_wall_ms": index_wall_ms,
            "e2e": {"value": bases_all * K / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mbp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / K, "device_ms_per_step": e2e_dev_ms / K,
                    "timed": "host clock around Engine.align (mfsc_align_batch + row collection) from pinned host buffers, max over ranks"},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"kernel": "k_extend<false,true>", "bound": "alu", "achieved": ext_tiops, "peak": alu_peak, "unit": "Tiop/s (int32)",
                         "frac": ext_tiops / alu_peak if alu_peak else None, "ops_per_cell": OPS_PER_CELL,
                         "traffic": traffic_of("k_extend"),
                         "issue": {"warp_instructions": ext_tr.get("warp_instructions"),
                                   "frac_of_issue_slots": (ext_tr["warp_instructions"] / (ext_tr["gpu_time_ms"] * 1e-3) / (n_sm * 4 * 1965e6))
                                   if ext_tr.get("warp_instructions") else None,
                                   "thread_instructions_per_cell": ext_tr.get("thread_instructions_per_cell"),
                                   "alu_pipe_active_pct": ext_tr.get("alu_pipe_active_pct"), "source": ext_tr.get("source")},
                         "note": "banded integer DP, no HBM traffic to speak of: achieved = cells x %d int ops (the SASS count of one packed cell: "
                                 "1 VIMNMX3 + 2 VIMNMX + 3 LOP3 + 5 IMAD + 1 predicated IMAD) / kernel time; peak = %d SMs x 128 lanes x %.0f MHz "
                                 "observed.  `issue` is the second view: what the kernel really issues (ncu), bookkeeping included"
                                 % (OPS_PER_CELL, n_sm, sm_mhz)},
            "roofline_seed": hbm_line("k_seeds", seed_bytes, seed_ms, "stage-2 time includes the match-count readback; algorithmic bytes = Q/2 + 8 P + 4 M "
                                      "(Q query bases of both strands, P probes, M matches)"),
            "roofline_cluster": hbm_line("k_cluster", 16 * M + 12 * M, clu_ms, "16 B per match read + 12 B per cluster match written; the kernel is bound "
                                         "by the serial mgaps rules per strand (lanes active, see profiles/), not by bytes"),
            "roofline_index": hbm_line("k_index_*", R_bases // 4 * 2 + 4 * info["positions"] + 8 * (4 ** info["k"]), tables_ms,
                                       "stage 1 tables: 0.25 R read twice (count, fill) + 4 R positions written + 8 x 4^k for the dense scratch histogram "
                                       "(zero + scan) that the compact directory + heads are made from"),
            "roofline_coverage": hbm_line("k_cov_*", cov_bytes, cov_ms, "stage 5 over this step's rows: 8 N_rows + 12 C (difference array written, "
                                          "read, profile written)") if cov_ms else None,
            "strong": strong,
        }
        if args.scaling == "strong" and strong:
            out.update({"value": strong["value"], "ms_per_step": strong["ms_per_step"], "scaling": "strong", "steps": strong["steps"],
                        "config": {"workload": strong["workload"], "parallelism": "one job, read batches over %d rank(s)" % world},
                        "weak_cfg": {"value": bases_all * K / (dev_ms_max * 1e-3) / 1e6, "workload": label}})
        if world == 1 and not args.no_wrapper and args.workload in ("cfg1", "cfg2"):
            out["e2e_wrapper"] = measure_wrapper(contigs, reads, n_bases)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import mummer_probe
            from oracle import oracle as O
            threads = os.cpu_count() or 1
            rl, cl = as_list(reads), as_list(contigs)
            ref = O.Ref(cl)
            n_s = size_sample(ref, rl, threads, args.cpu_seconds)
            orows, dt, ost = oracle_run(ref, rl[:n_s], threads)
            sample_bases = sum(len(x) for x in rl[:n_s])
            out["cpu_baseline"] = {"value": sample_bases / dt / 1e6, "unit": "Mbp/s", "cores": threads, "kind": "port",
                                   "sample": "first %d reads of the batch (%.1f s), CPU oracle restatement (MUMmer not in image), reference "
                                             "indexed once and shared by the threads" % (n_s, dt),
                                   "dp_gcups": ost["cells"] / dt / 1e9}
            # parity on the bench batch itself: the GPU rows of the sampled reads against the oracle's
            g = rows_matrix(r)
            g = g[g[:, 1] < n_s]
            equal = bool(g.shape == orows.shape and np.array_equal(g, orows))
            out["parity"] = {"reads": n_s, "rows": int(len(orows)), "rows_gpu": int(len(g)), "equal": equal,
                             "fields": "ref, read, strand, S1, E1, S2, E2, errors, columns (delta order)",
                             "oracle_max_band": ost["max_band"], "oracle_cap_hits": ost["cap_hits"]}
            out["mummer"] = mummer_probe.report()
            if not equal:
                emit(json.dumps(out))
                raise SystemExit("parity: GPU rows differ from the oracle's on the bench batch")
        emit(json.dumps(out))
    if comm is not None:
        comm.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
