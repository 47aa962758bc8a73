"""Run-time probe for a real MUMmer 3.23 (TEST INFRASTRUCTURE ONLY).

MUMmer is not part of the reference tree (README.md:41-45 names the download), of this image or -- as far as could be
seen -- of the GPU boxes, which is why the oracle is "parity unpinned" against it.  Every bench run and one test still
look for it where the reference itself expects it: `$PATH`, `/tmp/MUMmer3.23/` (onlineTest.py:14, .travis.yml:11-13),
`baseline/_ref/` of this repository, `$MUMMER_DIR`.  When it is found, `diff_with_oracle` runs the reference's own
two commands (alignerRobot.py:46-72: `nucmer --maxmatch -p <prefix> <ref> <qry>`, `show-coords -r <prefix>.delta`) on a
small seeded case and compares the coords text with the oracle's, and the result is reported in the bench line."""
import os
import shutil
import subprocess
import tempfile

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEARCHED = ["$PATH", "/tmp/MUMmer3.23/", os.path.join(_ROOT, "baseline", "_ref"), "$MUMMER_DIR"]


def find_mummer():
    """Directory holding executable `nucmer` and `show-coords`, or None."""
    cands = []
    w = shutil.which("nucmer")
    if w:
        cands.append(os.path.dirname(w))
    cands += ["/tmp/MUMmer3.23", os.path.join(_ROOT, "baseline", "_ref"), os.path.join(_ROOT, "baseline", "_ref", "MUMmer3.23"),
              os.path.join(_ROOT, "baseline", "_ref", "bin")]
    if os.environ.get("MUMMER_DIR"):
        cands.append(os.environ["MUMMER_DIR"])
    for d in cands:
        if all(os.path.isfile(os.path.join(d, x)) and os.access(os.path.join(d, x), os.X_OK) for x in ("nucmer", "show-coords")):
            return d
    return None


def run_reference_commands(mummer_dir, ref_fasta, qry_fasta, prefix, extra=""):
    """The two shell commands of alignerRobot.py:46-72; returns the coords text."""
    subprocess.check_call("%s/nucmer %s --maxmatch -p %s %s %s" % (mummer_dir, extra, prefix, ref_fasta, qry_fasta), shell=True,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return subprocess.check_output("%s/show-coords -r %s.delta" % (mummer_dir, prefix), shell=True).decode()


def diff_with_oracle(mummer_dir, n_reads=60):
    """Seeded small case (README shape): rows of real nucmer + show-coords against the oracle's.  Returns a dict for the
    bench line: rows of either side, equal coordinate rows, largest %IDY difference."""
    from metafinishersc_b200 import datagen, fasta
    from oracle import oracle as O
    d = datagen.abun_gap(G=100_000, repeat_len=1200, L=3000, seed=3)
    reads = d["reads"][:n_reads]
    with tempfile.TemporaryDirectory() as tmp:
        rn, qn = datagen.names("Contig", len(d["contigs"])), datagen.names("Segkk", len(reads))
        fasta.write_fasta(os.path.join(tmp, "ref.fasta"), rn, d["contigs"])
        fasta.write_fasta(os.path.join(tmp, "qry.fasta"), qn, reads)
        text = run_reference_commands(mummer_dir, os.path.join(tmp, "ref.fasta"), os.path.join(tmp, "qry.fasta"), os.path.join(tmp, "out"))
    real = []
    for ln in text.split("\n")[5:]:
        if not ln.strip():
            continue
        f = ln.split("|")
        a, b, c = f[0].split(), f[1].split(), f[2].split()
        names = f[-1].split("\t")
        real.append((int(a[0]), int(a[1]), int(b[0]), int(b[1]), int(c[0]), int(c[1]), float(f[3]), names[0].strip(), names[1].strip()))
    o = O.align(d["contigs"], reads, band_cap=1 << 20)
    mine = [tuple(r) for r in O.coords_rows(o["rows"], rn, qn)]
    key = lambda r: (r[7], r[8], r[0], r[1], r[2], r[3])
    rs, ms = sorted(real, key=key), sorted(mine, key=key)
    same = sum(1 for x, y in zip(rs, ms) if key(x) == key(y)) if len(rs) == len(ms) else len(set(map(key, rs)) & set(map(key, ms)))
    didy = max([abs(x[6] - y[6]) for x, y in zip(rs, ms) if key(x) == key(y)] or [0.0])
    return {"found": True, "dir": mummer_dir, "rows_mummer": len(rs), "rows_oracle": len(ms), "rows_equal": same,
            "max_idy_diff": didy, "equal": len(rs) == len(ms) == same and didy <= 0.01}


def report():
    d = find_mummer()
    if d is None:
        return {"found": False, "searched": SEARCHED}
    try:
        return diff_with_oracle(d)
    except Exception as e:       # a broken installation must not take the bench down
        return {"found": True, "dir": d, "error": str(e)[:200]}
