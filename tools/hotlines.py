#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions of one kernel: joins `ncu --page source --csv` (SASS view)
with the line table of the cubin (`nvdisasm -g`).  usage: tools/hotlines.py <report.ncu-rep> <kernel regex> <mangled-name-substring> [top]"""
import csv
import re
import subprocess
import sys
import tempfile

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = __import__("os").environ.get("HOTLINES_SO", "metafinishersc_b200/csrc/libmfsc_b200.so")    # the build that was profiled
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", __import__("os").path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = [f for f in __import__("os").listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", tmp + "/" + cub], capture_output=True, text=True).stdout.split("\n")
# line table: instruction index within the function -> (file, line)
inside, cur, lines = False, ("?", 0), []
for ln in dis:
    if ln.startswith(".text.") and mangled in ln and ln.rstrip().endswith(":"):
        inside = True
        continue
    if inside:
        if ln.startswith("//-----") or ln.startswith("\t.section"):
            if lines:
                break
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
H = rows[1]
si, ii, ti = H.index("# Samples"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
agg = {}
k = 0
body = [r for r in rows[2:] if len(r) > ti and r[si].isdigit()]
if len(body) == 2 * len(lines):          # ncu lists the function twice in this view
    body = body[:len(lines)]
for r in body:
    key = lines[k] if k < len(lines) else ("?", -1)
    k += 1
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[ti])
ts, tins = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("instructions in function: sass %d, line table %d; samples %d, warp instructions %d" % (k, len(lines), ts, tins))
src = {}
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open("metafinishersc_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
    print("%5.1f%% smp %5.1f%% ins %4.1f lanes | %s:%d | %s" % (100.0 * a[0] / max(ts, 1), 100.0 * a[1] / max(tins, 1), a[2] / max(a[1], 1), f, l, text))
