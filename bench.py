#!/usr/bin/env python
"""bench.py -- read->contig alignment throughput (aligned Mbp/s) of the hot path on N B200s.

A step = one pass of stages 2-4 (seeds, clusters, extension) over one batch of synthetic reads
against a contig index resident in HBM.  Workload at every N: BASELINE.json configs[1] -- one
5 Mbp genome, 50X of 10 kbp reads with 15 % indel error (25 000 reads, ~250 Mbp) PER GPU (weak
scaling: reads shard by batch, no data-path collective; the index is built on rank 0 and
replicated with one NCCL broadcast per buffer).

  value : Mbp/s with the read batch already packed in HBM (CUDA events around the step)
  e2e   : same metric through mfsc_align_batch with pinned HOST buffers (H2D of the reads and D2H
          of the rows inside the timed region)
  --impl reference : the CPU oracle (restatement of nucmer+show-coords; MUMmer itself is not in the
          image) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from metafinishersc_b200 import datagen  # noqa: E402

METRIC = "read->contig aligned Mbp/s (nucmer --maxmatch path: seeds+clusters+extension)"
WORKLOAD = "cfg2: 5 Mbp genome, 50X, 10 kbp reads, 15% indel (p_del=p_ins=0.075), 25000 reads/GPU"


def make_workload(rank, n_reads, genome, workload="cfg2"):
    """Every rank aligns against the same contigs; the reads of rank r are drawn with their own seed."""
    if workload == "cfg1":      # README shape: two 5 Mbp genomes at 50X/20X, 6 kbp reads at 1% del + 1% ins, 2 chimeric contigs
        d = datagen.abun_gap(G=genome, repeat_len=12000, L=6000, p=0.01, abun=(50, 20), seed=rank)
        g = datagen.abun_gap(G=genome, repeat_len=12000, L=6000, p=0.01, abun=(0.01, 0.01), seed=0) if rank else d
        return g["contigs"], d["reads"][:n_reads] if n_reads < len(d["reads"]) else d["reads"]
    d = datagen.sc_lr(G=genome, n_reads=n_reads, seed=100, read_seed=2100 + rank)
    return d["contigs"], d["reads"]


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


def oracle_rate(contigs, reads, threads):
    """Mbp/s of the CPU oracle over `reads` on `threads` host threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    chunks = [reads[i::threads] for i in range(threads)]
    chunks = [c for c in chunks if c]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda c: O.align(contigs, c, band_cap=248)["stats"]["cells"], chunks))
    dt = time.perf_counter() - t0
    bases = sum(len(r) for r in reads)
    return bases / dt / 1e6, dt, sum(res)


def size_sample(contigs, reads, threads, target_s):
    """Reads per sample so that the oracle runs about target_s on `threads` threads: two pilots fit
    time = fixed (per-thread index build) + per-read cost."""
    n1, n2 = 2 * threads, 8 * threads
    _, t1, _ = oracle_rate(contigs, reads[:n1], threads)
    _, t2, _ = oracle_rate(contigs, reads[:n2], threads)
    b = max((t2 - t1) / (n2 - n1), 1e-6)
    a = max(t1 - b * n1, 0.0)
    return int(max(threads, min(len(reads), (target_s - a) / b)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    contigs, reads = make_workload(0, args.reads, args.genome)
    # pilots size the per-step sample: ~ args.ref_step_seconds of CPU work on all threads, less when many steps are asked
    # for, so that the whole run stays within a few minutes
    step_s = min(args.ref_step_seconds, 150.0 / max(1, args.steps + args.warmup))
    n_sample = size_sample(contigs, reads, threads, step_s)
    reads = reads[:n_sample]
    for _ in range(args.warmup):
        oracle_rate(contigs, reads[:threads], threads)
    total_dt, total_bases, cells = 0.0, 0, 0
    for _ in range(args.steps):
        r, dt, c = oracle_rate(contigs, reads, threads)
        total_dt += dt; total_bases += sum(len(x) for x in reads); cells += c
    v = total_bases / total_dt / 1e6
    sample = "%d reads (%.1f Mbp) per step of the 25000-read batch, CPU oracle (restatement, not MUMmer)" % (n_sample, total_bases / args.steps / 1e6)
    emit(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": total_dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "sample": sample},
                      "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": threads, "kind": "port", "sample": sample,
                                       "dp_gcups": cells / total_dt / 1e9},
                      "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (NCCL's
    version banner, library chatter) has been diverted to stderr by main()."""
    os.write(_STDOUT_FD, (line + "\n").encode())


_STDOUT_FD = 1


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
    ap.add_argument("--reads", type=int, default=25000, help="reads per GPU per step (cfg2: 25000)")
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--ref-step-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", type=int, default=0, help="stage-4 DP engine: 0 packed (+ general for handed-over reads), 1 general only")
    ap.add_argument("--kmer", type=int, default=0, help="index k-mer length (0: chosen from the reference size)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg1"], help="cfg2 is the headline; cfg1 = README shape (parity/aux case)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    from metafinishersc_b200 import engine, multigpu
    torch.cuda.set_device(local)
    E = engine.Engine()
    E.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = engine.make_params(engine=args.engine, kmer=args.kmer)
    contigs, reads = make_workload(rank, args.reads if args.workload == "cfg2" else 10**9, args.genome, args.workload)
    buf, off = engine._as_buf(reads)
    n_bases = int(off[-1])

    # stage 1 on rank 0, replicated with one NCCL broadcast per buffer
    t0 = time.perf_counter()
    ix, bcast_ms = multigpu.replicated_index(E, contigs, P, rank, world)
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
        return E.align(ix, (hbuf, hoff), P)

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
    e2e_ms = 0.0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        re = step_e2e()
        e2e_ms += re.ms["total"]
    barrier()
    clk = clocks.stop() if rank == 0 else None
    h2d, d2h = re.stats["h2d_bytes"], re.stats["d2h_bytes"]

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_ms_max = max_over_ranks(dev_ms)
    e2e_ms_max = max_over_ranks(e2e_ms)
    wall_ms_max = max_over_ranks(wall_ms)
    bases_all = sum_over_ranks(float(n_bases))
    cells_all = sum_over_ranks(float(cells))
    ext_ms_max = max_over_ranks(stage["extend"])      # every collective happens on all ranks, before the rank-0 block
    K = args.steps
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except Exception:
            pass
        full = args.workload == "cfg2" and args.reads == 25000 and args.genome == 5_000_000
        ext_tr = traffic.get("k_extend", {}) if full else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        ext_ms = stage["extend"] / K
        seed_ms = stage["seeds"] / K
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        # Algorithmic work of one DP cell: the three-state affine-gap recurrence with the two carried counters,
        # written out on separate words, is 40 integer operations (the count the first version of the kernel issued).  The
        # packed engine issues about half of that per cell; the figure is kept so that the fraction is comparable over time.
        ops_per_cell = 40
        alu_peak = n_sm * 128 * sm_mhz * 1e6 / 1e12   # Tiop/s at the clock observed in this run
        ext_tiops = cells * ops_per_cell / (ext_ms * 1e-3) / 1e12 if ext_ms > 0 else 0.0
        out = {
            "metric": METRIC, "value": bases_all * K / (dev_ms_max * 1e-3) / 1e6, "unit": "Mbp/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "cfg1: README shape, 2 x 5 Mbp contigs, %d reads x 6 kbp, 1%% del + 1%% ins" % len(reads) if args.workload == "cfg1" else
                       WORKLOAD if args.reads == 25000 and args.genome == 5_000_000 else
                       "cfg2 shape scaled: %d bp genome, %d reads/GPU" % (args.genome, args.reads),
                       "reads_per_gpu": args.reads, "query_bases_per_gpu": n_bases, "l2": "flushed (256 MiB write) between steps",
                       "dp_engine": "packed 32-bit states, general engine for handed-over reads" if args.engine == 0 else "general",
                       "index_k": info["k"], "index_hbm_bytes": info["hbm_bytes"], "parallelism": "reads sharded by batch x%d" % world},
            "wall_ms_per_step": wall_ms_max / K,
            "stage_ms_per_step": {k: v / K for k, v in stage.items()},
            "dp_gcups": cells_all / (ext_ms_max / K * 1e-3) / 1e9 if ext_ms_max > 0 else None,
            "dp_cells_per_step": cells_all, "rows_per_step_rank0": n_rows, "dp_reads_handed_over_rank0": r.stats.get("dp_retry_reads"),
            "index_build_ms": ix.build_ms() if rank == 0 else None, "index_broadcast_ms": bcast_ms, "index_wall_ms": index_wall_ms,
            "e2e": {"value": bases_all * K / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mbp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / K},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"kernel": "k_extend", "bound": "alu", "achieved": ext_tiops, "peak": alu_peak, "unit": "Tiop/s (int32)",
                         "frac": ext_tiops / alu_peak if alu_peak else None,
                         "traffic": (ext_tr["dram_bytes_read"] + ext_tr["dram_bytes_write"]) if ext_tr else None,
                         "pipe_active_ncu": {"alu": ext_tr.get("alu_pipe_active_pct"), "fma": ext_tr.get("fma_pipe_active_pct"),
                                             "issue": ext_tr.get("issue_active_pct"), "source": ext_tr.get("source")} if ext_tr else None,
                         "note": "banded DP, no HBM traffic to speak of (traffic = ncu dram bytes per launch): cells x %d algorithmic int ops / "
                                 "kernel time; peak = %d SMs x 128 lanes x %.0f MHz observed. What binds is instruction issue: the ncu capture "
                                 "(pipe_active_ncu) shows the issue slots and the ALU pipe (64 lanes/clk/SM) about 80 %% busy"
                                 % (ops_per_cell, n_sm, sm_mhz)},
            "roofline_seed": {"kernel": "k_seeds", "bound": "hbm", "achieved": seed_bytes / (seed_ms * 1e-3) / 1e9 if seed_ms > 0 else None,
                              "peak": hbm_peak, "unit": "GB/s", "frac": seed_bytes / (seed_ms * 1e-3) / 1e9 / hbm_peak if seed_ms > 0 else None,
                              "traffic": None, "peak_source": peak_src,
                              "note": "stage-2 time includes the match-count readback; algorithmic bytes = Q/2 + 8 P + 4 M"},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_s = size_sample(contigs, reads, threads, args.cpu_seconds)
            v, dt, c = oracle_rate(contigs, reads[:n_s], threads)
            out["cpu_baseline"] = {"value": v, "unit": "Mbp/s", "cores": threads, "kind": "port",
                                   "sample": "first %d reads of the batch (%.1f s), CPU oracle restatement (MUMmer not in image)" % (n_s, dt),
                                   "dp_gcups": c / dt / 1e9}
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
