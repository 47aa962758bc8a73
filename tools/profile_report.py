#!/usr/bin/env python
"""Turns one ncu report (several kernels, `--set full`) into the tracked summaries: profiles/<tag>_<kernel>_ncu.md per kernel and the
entries of profiles/r02_traffic.json that bench.py reads (`traffic` of the roofline lines).
usage: tools/profile_report.py <tag> <workload cfg1|cfg2|...> <report.ncu-rep> "<command that was profiled>" [cells]"""
import csv
import json
import os
import subprocess
import sys

tag, workload, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
cells = float(sys.argv[5]) if len(sys.argv) > 5 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def num(v, unit):
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit)
    return x * scale if scale else x


tpath = "profiles/r02_traffic.json"
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
for r in rows[2:]:
    if len(r) < len(H):
        continue
    name = r[H.index("Kernel Name")].split("(")[0].strip()
    short = name.replace("void ", "").split("<")[0]
    m = {h: (r[i], U[i]) for i, h in enumerate(H)}
    path = "profiles/%s_%s_ncu.md" % (tag, short)
    with open(path, "w") as f:
        f.write("# %s -- ncu --set full of `%s` (one launch, %s)\n\nCommand: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (tag, name, workload, cmd))
        for w in WANT:
            if w in m:
                f.write("| %s | %s | %s |\n" % (w, m[w][0], m[w][1]))
    e = {"workload": workload, "dram_bytes_read": num(*m["dram__bytes_read.sum"]), "dram_bytes_write": num(*m["dram__bytes_write.sum"]),
         "gpu_time_ms": num(*m["gpu__time_duration.sum"]), "warp_instructions": float(m["smsp__inst_executed.sum"][0].replace(",", "")),
         "lanes_active": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
         "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
         "alu_pipe_active_pct": float(m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", ("nan", ""))[0]),
         "l2_hit_pct": float(m["lts__t_sector_hit_rate.pct"][0]), "source": path}
    if cells and short == "k_extend":
        e["thread_instructions_per_cell"] = e["warp_instructions"] * e["lanes_active"] / cells
        e["warp_instructions_per_cell"] = e["warp_instructions"] / cells
    traffic.setdefault(workload, {})[short] = e
    print(short, json.dumps(e))
json.dump(traffic, open(tpath, "w"), indent=1)
