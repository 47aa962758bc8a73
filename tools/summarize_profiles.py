#!/usr/bin/env python
"""Turns the ncu outputs that gpurun brought back (gpurun_out/) into the tracked summaries under profiles/.
usage: tools/summarize_profiles.py <tag> <launches.csv> <report.ncu-rep> [kernel-label]"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    return {h: (V[i], U[i]) for i, h in enumerate(H)}


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    tag, lcsv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    label = sys.argv[4] if len(sys.argv) > 4 else "k_extend"
    cmd = sys.argv[5] if len(sys.argv) > 5 else ""
    agg = launches(lcsv)
    tot = sum(v for _, v in agg.values())
    with open("profiles/%s_launches.md" % tag, "w") as f:
        f.write("# %s -- ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
        f.write("Command: `%s`. Cold-cache, serialised launch times: compare SHARES, not absolutes.\n\n" % cmd)
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f%% |\n" % (k, n, v, 100 * v / tot))
    m = raw_metrics(rep)
    with open("profiles/%s_%s_ncu.md" % (tag, label.replace("k_", "")), "w") as f:
        f.write("# %s -- ncu --set full of `%s` (one launch)\n\nCommand: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (tag, label, cmd))
        for w in WANT:
            if w in m:
                f.write("| %s | %s | %s |\n" % (w, m[w][0], m[w][1]))


if __name__ == "__main__":
    main()
