"""Generates the committed golden fixtures of tests/golden/ (run in the authoring container, where
/root/reference is mounted; the fixtures travel, the reference does not).

1. fasta_splitter.json -- record ranges written by the reference's own Perl splitter
   (/root/reference/srcRefactor/repeatPhaserLib/fasta-splitter.pl --n-parts N), run here under the
   system perl on seeded toy FASTA files.  Pins metafinishersc_b200.fasta.part_boundaries / split_fasta.
2. readme_tables.json -- the show-coords tables printed in the reference README.md:98-129 (literal
   known answers; tables 2 and 3 are structure-determined and reproducible offline) and the row
   layout sample of toCondenseFixer.py:133-139.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SPLITTER = "/root/reference/srcRefactor/repeatPhaserLib/fasta-splitter.pl"


def toy_fasta(seed, n, lo, hi, wrap):
    rng = np.random.default_rng(seed)
    names = ["Segkk%d" % i for i in range(n)]
    seqs = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(lo, hi)))) for _ in range(n)]
    lines = []
    for nm, s in zip(names, seqs):
        lines.append(">" + nm)
        if wrap:
            lines += [s[i:i + wrap] for i in range(0, len(s), wrap)]
        else:
            lines.append(s)
    return names, seqs, "\n".join(lines) + "\n"


def run_splitter(text, n_parts):
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "reads.fasta"), "w") as f:
            f.write(text)
        subprocess.check_call(["perl", SPLITTER, "--n-parts", str(n_parts), os.path.join(d, "reads.fasta")], cwd=d,
                              stdout=subprocess.DEVNULL)
        parts = sorted(p for p in os.listdir(d) if ".part-" in p)
        out = []
        for p in parts:
            body = open(os.path.join(d, p)).read()
            out.append({"file": p, "records": [ln[1:] for ln in body.split("\n") if ln.startswith(">")],
                        "sha1": hashlib.sha1(body.encode()).hexdigest()})
        return out


def main():
    cases = []
    for seed, n, lo, hi, wrap, n_parts in [(1, 50, 50, 400, 0, 20), (2, 7, 10, 30, 0, 20), (3, 200, 1, 120, 70, 20),
                                           (4, 33, 500, 900, 60, 4), (5, 20, 100, 101, 0, 20), (6, 1, 100, 200, 0, 3)]:
        names, seqs, text = toy_fasta(seed, n, lo, hi, wrap)
        cases.append({"seed": seed, "n": n, "lo": lo, "hi": hi, "wrap": wrap, "n_parts": n_parts, "parts": run_splitter(text, n_parts)})
    json.dump(cases, open(os.path.join(HERE, "fasta_splitter.json"), "w"), indent=1)
    readme = {
        "source": "reference README.md:98-129 (show-coords without -r: rows in delta order) and toCondenseFixer.py:133-139",
        "abun_vs_reference": ["2500006  2512005  |  2500001  2512000  |    12000    12000  |   100.00  | Segkk1\tSegkk0",
                              "      1  5000002  |  5000000        1  |  5000002  5000000  |    99.99  | Segkk0\tSegkk0",
                              "      1  5000005  |        1  5000000  |  5000005  5000000  |    99.99  | Segkk1\tSegkk1",
                              "2488001  2500002  |  2512000  2500001  |    12002    12000  |    99.83  | Segkk0\tSegkk1"],
        "LC_vs_reference": [[1, 2512000, 1, 2512000, 2512000, 2512000, 100.00, "Segkk0", "Segkk0"],
                            [2500001, 5000000, 2500001, 5000000, 2500000, 2500000, 100.00, "Segkk1", "Segkk0"],
                            [1, 2512000, 1, 2512000, 2512000, 2512000, 100.00, "Segkk1", "Segkk1"],
                            [2500001, 5000000, 2500001, 5000000, 2500000, 2500000, 100.00, "Segkk0", "Segkk1"]],
        "reference_vs_reference": [[1, 5000000, 1, 5000000, 5000000, 5000000, 100.00, "Segkk0", "Segkk0"],
                                   [2500001, 2512000, 2500001, 2512000, 12000, 12000, 100.00, "Segkk1", "Segkk0"],
                                   [1, 5000000, 1, 5000000, 5000000, 5000000, 100.00, "Segkk1", "Segkk1"],
                                   [2500001, 2512000, 2500001, 2512000, 12000, 12000, 100.00, "Segkk0", "Segkk1"]],
        "row_layout_samples": ["       1      562  |      819     1418  |      562      600  |    84.72  | Contig0_d\tRead121_d",
                               "       1      562  |     4077     3478  |      562      600  |    84.72  | Contig0_d\tRead121_p",
                               "       1      564  |      656       68  |      564      589  |    90.13  | Contig0_d\tRead382_d"],
        "row_layout_parsed": [[1, 562, 819, 1418, 562, 600, 84.72, "Contig0_d", "Read121_d"],
                              [1, 562, 4077, 3478, 562, 600, 84.72, "Contig0_d", "Read121_p"],
                              [1, 564, 656, 68, 564, 589, 90.13, "Contig0_d", "Read382_d"]],
    }
    json.dump(readme, open(os.path.join(HERE, "readme_tables.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
