"""FASTA ingest and the read-part splitter of the alignment path.

`split_fasta` restates `fasta-splitter.pl --n-parts N` (reference
srcRefactor/repeatPhaserLib/fasta-splitter.pl:135-183): greedy split on measure=all byte
size, records re-wrapped at 60 columns, parts named <base>.part-NN.fasta.  The part
boundaries define the read batches (and therefore the row order) that
alignerRobot.largeRvsQAlign (alignerRobot.py:258-297) concatenates.
"""
import os

import numpy as np


def read_fasta(path):
    """-> (names list[str], seqs list[bytes]).  Name = header text after '>' (IORobot.py:21)."""
    names, seqs, cur = [], [], []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line[:1] == b">":
                if names:
                    seqs.append(b"".join(cur))
                names.append(line[1:].decode())
                cur = []
            else:
                cur.append(line)
    if names:
        seqs.append(b"".join(cur))
    return names, seqs


def read_fasta_arrays(path):
    """Vectorised ingest for the aligner: -> (names list[str], uint8 bases, int64 offsets[n+1]).
    Same records as read_fasta (blank lines skipped, sequence lines joined), without per-line Python work."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size == 0:
        return [], np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    nl = np.flatnonzero(data == 10)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl, [data.size]))
    if starts[-1] >= data.size:                       # file ends with a newline
        starts, ends = starts[:-1], ends[:-1]
    ends = ends - ((ends > starts) & (data[np.maximum(ends - 1, 0)] == 13))      # strip \r
    nonempty = ends > starts
    starts, ends = starts[nonempty], ends[nonempty]
    is_hdr = data[starts] == 62
    hdr_idx = np.flatnonzero(is_hdr)
    if hdr_idx.size == 0:
        return [], np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    raw = data.tobytes()
    names = [raw[starts[i] + 1:ends[i]].decode() for i in hdr_idx]
    # sequence lines before the first header are ignored (read_fasta does the same)
    seq_mask = ~is_hdr
    seq_mask[:hdr_idx[0]] = False
    lens = np.where(seq_mask, ends - starts, 0)
    # record of each line = number of headers seen so far - 1
    rec_of_line = np.cumsum(is_hdr) - 1
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.add.at(off, rec_of_line[seq_mask] + 1, lens[seq_mask])
    np.cumsum(off, out=off)
    # gather the bases: byte-level keep mask from a +1/-1 difference array over line spans
    diff = np.zeros(data.size + 1, dtype=np.int8)
    np.add.at(diff, starts[seq_mask], 1)
    np.add.at(diff, ends[seq_mask], -1)
    keep = np.cumsum(diff[:-1], dtype=np.int8).astype(bool)
    return names, np.ascontiguousarray(data[keep]), off


def write_fasta(path, names, seqs, line_len=0):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            if line_len:
                for i in range(0, len(s), line_len):
                    f.write(s[i:i + line_len] + b"\n")
            else:
                f.write(s + b"\n")


def concat(seqs):
    """list[bytes] -> (uint8 array of all bases, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=off[1:])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return buf, off


def obtain_length(path):
    """IORobot.obtainLength (IORobot.py:8-30): name -> summed sequence-line length."""
    names, seqs = read_fasta(path)
    return {n: len(s) for n, s in zip(names, seqs)}


def part_boundaries(names, seqs, n_parts, line_len=60):
    """Record index ranges [(begin, end), ...] of the parts fasta-splitter.pl would write
    (fasta-splitter.pl:160-183, measure=all, unix eol)."""
    def seq_size(name, s):
        slen = len(s)
        return slen + len(name) + 1 + (1 + ((slen + line_len - 1) // line_len if line_len else 1))
    sizes = [seq_size(n, s) for n, s in zip(names, seqs)]
    total = sum(sizes)
    parts, begin = [], 0
    written_total, written_this, part, part_end, opened = 0, 0, 1, int(total / n_parts), False
    for i, sz in enumerate(sizes):
        new_total = written_total + sz
        if (not opened) or (written_this and new_total > part_end and new_total - part_end > part_end - written_total):
            if opened:
                parts.append((begin, i))
                begin = i
                part += 1
            opened = True
            part_end = int(total / n_parts * part) + 1
            written_this = 0
        written_this += sz
        written_total += sz
    if opened:
        parts.append((begin, len(sizes)))
    return parts


def split_fasta(path, n_parts, out_dir=None, line_len=60):
    """Write <base>.part-NN.fasta files; returns their paths.  fasta-splitter.pl writes into
    the CWD and alignerRobot.py:268-269 then copies them into the folder; callers pass
    out_dir=folder to land them there directly."""
    names, seqs = read_fasta(path)
    base = os.path.basename(path)
    ext = ""
    for e in (".fasta", ".faa", ".fna", ".fa"):
        if base.lower().endswith(e) and len(base) > len(e):
            base, ext = base[:-len(e)], base[-len(e):]
            break
    out_dir = out_dir if out_dir is not None else os.getcwd()
    num_len = len(str(n_parts))
    paths = []
    for p, (b, e) in enumerate(part_boundaries(names, seqs, n_parts, line_len), start=1):
        stem = base if (".part-" in base and base.rsplit(".part-", 1)[1].isdigit()) else base + ".part"
        out = os.path.join(out_dir, "%s-%0*d%s" % (stem, num_len, p, ext))
        write_fasta(out, names[b:e], seqs[b:e], line_len)
        paths.append(out)
    return paths
