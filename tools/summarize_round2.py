#!/usr/bin/env python
"""Builds profiles/r02_summary.md from the bench lines kept under profiles/r02_bench/ (one JSON line per file, copied from
gpurun_out/ by hand) and profiles/r02_traffic.json (tools/profile_report.py)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "profiles", "r02_bench")


def load(name):
    p = os.path.join(B, name)
    if not os.path.exists(p):
        return None
    try:
        return json.loads(open(p).read().strip().split("\n")[-1])
    except Exception:
        return None


def f(x, nd=1):
    return "—" if x is None else ("%." + str(nd) + "f") % x


out = ["# r02 — numbers of this round (one B200 unless stated; SM clock 1965 MHz, no throttle reasons in any run)", "",
       "Every line below is a `bench.py` JSON line kept under `profiles/r02_bench/` (the command is in the file name / `config`);",
       "ncu summaries are the `r02_*_ncu.md` files beside this one, `r02_traffic.json` feeds `bench.py`'s `traffic` fields.", ""]
d = load("default_n1.json")
if d:
    out += ["## Default run (`python bench.py`), N = 1", "",
            "| quantity | value |", "|---|---|",
            "| `value` (cfg 2, batch resident in HBM) | %s Mbp/s, %s ms per 250 Mbp step |" % (f(d["value"], 0), f(d["ms_per_step"])),
            "| `e2e` (host clock around `Engine.align`, pinned host buffers) | %s Mbp/s (%s ms; H2D %d B, D2H %d B) |" % (f(d["e2e"]["value"], 0), f(d["e2e"]["ms_per_step"]), d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"]),
            "| `e2e_wrapper` (`alignerRobot.largeRvsQAlign`, files in, files + rows out) | %s Mbp/s (%s s); with MUMmer delta records %s Mbp/s |" % (f(d["e2e_wrapper"]["value"], 0), f(d["e2e_wrapper"]["seconds"], 3), f(d["e2e_wrapper"]["with_full_delta_records"]["value"], 0)) if d.get("e2e_wrapper") else "| e2e_wrapper | — |",
            "| stages (ms per step) | %s |" % ", ".join("%s %s" % (k, f(v, 2)) for k, v in d["stage_ms_per_step"].items()),
            "| DP | %.3g cells per step, %s GCUPS; widest band %d, diagonals where `band_cap` bound: %d; reads handed to the general engine: %s |" % (d["dp_cells_per_step"], f(d["dp_gcups"], 0), d["max_band"], d["diagonals_where_the_cap_bound"], d["dp_reads_handed_over_rank0"]),
            "| parity on the bench batch | %s |" % json.dumps(d.get("parity")),
            "| CPU oracle, all host threads | %s |" % json.dumps({k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"}),
            "| MUMmer probe | %s |" % json.dumps(d.get("mummer")),
            ""]
    out += ["### Rooflines of the default run", "", "| stage | kernel | bound | achieved | peak | frac | algorithmic bytes | ncu DRAM bytes per launch |", "|---|---|---|---:|---:|---:|---:|---:|"]
    for key, stage in (("roofline_index", "1 index"), ("roofline_seed", "2 seeds"), ("roofline_cluster", "3 clusters"), ("roofline", "4 extension"), ("roofline_coverage", "5 coverage")):
        r = d.get(key)
        if not r:
            continue
        out.append("| %s | `%s` | %s | %s %s | %s | %s | %s | %s |" % (stage, r["kernel"], r["bound"], f(r["achieved"], 2), r["unit"].split(" ")[0], f(r["peak"], 1), f(r["frac"], 4),
                                                             r.get("algorithmic_bytes", "—"), r.get("traffic") if r.get("traffic") is not None else "—"))
    r = d["roofline"]
    out += ["", "`k_extend`: %d integer operations per cell (SASS); issue view from ncu: %s" % (r["ops_per_cell"], json.dumps(r.get("issue"))), ""]
out += ["## Workloads (BASELINE configs), N = 1, batch resident in HBM", "",
        "| workload | Mbp/s | ms per step | seeds | sort | cluster | extend | rows back | e2e Mbp/s | DP cells | index k / HBM bytes / tables ms | rows == CPU oracle on the sampled reads |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|"]
for name in ("cfg1.json", "default_n1.json", "cfg3.json", "cfg4.json", "cfg4_full.json", "cfg5.json"):
    d = load(name)
    if not d:
        continue
    s = d["stage_ms_per_step"]
    par = d.get("parity")
    out.append("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %.3g | %d / %d / %s | %s |" % (d["config"]["workload"], f(d["value"], 0), f(d["ms_per_step"], 2), f(s["seeds"], 2), f(s["sort"], 2), f(s["cluster"], 2),
                                                                                  f(s["extend"], 2), f(s["download"], 2), f(d["e2e"]["value"], 0), d["dp_cells_per_step"], d["config"]["index_k"],
                                                                                  d["config"]["index_hbm_bytes"], f(d.get("index_tables_ms"), 2),
                                                                                  ("%s: %d reads, %d rows" % ("equal" if par["equal"] else "DIFFERENT", par["reads"], par["rows"])) if par else "—"))
out += ["", "## 1 → 8 GPUs", "",
        "| N | weak: `value` Mbp/s (cfg 2 per GPU) | weak e2e | strong: cfg 3 as one job, Mbp/s | ms per step (mean / median) | efficiency vs N = 1 (mean / median) | rank-0 breakdown of the last step (ms) |", "|---:|---:|---:|---:|---:|---:|---|"]
base = None
for n, name in ((1, "default_n1.json"), (2, "n2.json"), (4, "n4.json"), (8, "n8.json")):
    d = load(name)
    if not d:
        continue
    s = d.get("strong") or {}
    med = s.get("ms_per_step_median") or s.get("ms_per_step")
    if n == 1 and s:
        base, base_med = s["ms_per_step"], med
    eff = base / (n * s["ms_per_step"]) if (s and base) else None
    eff_med = base_med / (n * med) if (s and base) else None
    bd = {k: round(v, 1) for k, v in (s.get("rank0_breakdown_ms") or {}).items() if isinstance(v, (int, float))}
    out.append("| %d | %s | %s | %s | %s / %s | %s / %s | %s |" % (n, f(d["value"], 0), f(d["e2e"]["value"], 0), f(s.get("value"), 0), f(s.get("ms_per_step"), 1), f(med, 1), f(eff, 3), f(eff_med, 3), json.dumps(bd)))
out += ["", "Limiting terms of the sharded job at N = 8 (18.8 ms per step against 77.5 ms on one GPU): (1) the reads are 1.5 GB of ASCII on the host and the",
        "ranks share the host's memory / PCIe bandwidth -- the rank's 187 MB are on the device after 12.4 ms (15 GB/s per rank, ~120 GB/s in aggregate;",
        "one GPU alone gets 43 GB/s), so the upload, which one GPU hides behind 65 ms of alignment, is longer than the alignment itself (9.3 ms of kernels); (2) the index:",
        "7.6 ms (contigs packed on rank 0 and broadcast 2.1, own 1/8 of the tables 1.4, exchange of the slices 4.0) against 11.2 ms on one GPU --",
        "it does not shrink because the exchange (three all-gathers of 64 + 180 + 200 MB and the wait for the slowest rank) replaces the build time it saves;",
        "(3) about a millisecond of fixed cost per GPU call (launches, stream synchronisations, sort set-up) and 1 ms for gathering the rows.",
        "Per-rank alignment itself scales: 65.5 ms / 8 = 8.2 ms against 9.3-9.7 ms measured.  `n4_strong_only_pieces*.json` (earlier build of this round):",
        "the same job at N = 4 run alone (`--strong-only`) with 1, 2 and 4 pieces per rank: 27.4 / 26.8 / 26.8 ms.", "",
        "Launch list of the default command (`profiles/r02_launches.csv`, `tools/launch_list.sh`: ncu `gpu__time_duration.sum`, cold-cache and serialised,",
        "shares only): `k_extend<0,1,0>` 97.5 % of the kernel time of a step, `k_cluster` 0.7 %, `k_seeds` 0.6 %, `k_pack` 0.6 %, the CUB radix sort 0.2 % --",
        "the same shares as the CUDA-event stage times of the bench line (extend 115.1 of 117.3 ms = 98.1 %).", ""]
tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
if os.path.exists(tp):
    t = json.load(open(tp))
    out += ["## ncu per kernel (one launch, `--set full`)", "", "| workload | kernel | ms | DRAM read + write (MB) | L2 hit % | lanes active | warp instructions | issue active % |", "|---|---|---:|---:|---:|---:|---:|---:|"]
    for w in sorted(t):
        for k, e in t[w].items():
            out.append("| %s | `%s` | %s | %s | %s | %s | %.4g | %s |" % (w, k, f(e["gpu_time_ms"], 3), f((e["dram_bytes_read"] + e["dram_bytes_write"]) / 1e6, 0), f(e["l2_hit_pct"]), f(e["lanes_active"], 2),
                                                                    e["warp_instructions"], f(e["issue_active_pct"])))
    out.append("")
open(os.path.join(ROOT, "profiles", "r02_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
