"""FASTA ingest and the read-part splitter of the alignment path.

`split_fasta` restates `fasta-splitter.pl --n-parts N` (reference
srcRefactor/repeatPhaserLib/fasta-splitter.pl:135-183): greedy split on measure=all byte
size, records re-wrapped at 60 columns, parts named <base>.part-NN.fasta.  The part
boundaries define the read batches (and therefore the row order) that
alignerRobot.largeRvsQAlign (alignerRobot.py:258-297) concatenates.
"""
import ctypes
import os
import threading
import weakref

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


# ------------------------------------------------------------------------------------------
# Page-locked host buffers for what is uploaded.  A copy from pageable memory is staged by the driver at a few GB/s (measured:
# 23 ms per 62 MB batch of parsed reads, most of it on the critical path of largeRvsQAlign); from page-locked memory the same
# copy runs at interconnect speed.  Page-locking is slow (tens of ms per 100 MB), so blocks are pooled: a block goes back to
# the pool when the last numpy view of it is collected, and the next parse of a similar size reuses it.
# ------------------------------------------------------------------------------------------
_PIN_LOCK = threading.Lock()
_PIN_FREE = []                    # [(bytes, address)] blocks not in use
_PIN_MIN, _PIN_MAX, _PIN_KEEP = 1 << 22, 1 << 31, 4     # below _PIN_MIN / above _PIN_MAX: pageable; blocks kept when free


def _pin_release(L, nbytes, addr):
    with _PIN_LOCK:
        if len(_PIN_FREE) < _PIN_KEEP:
            _PIN_FREE.append((nbytes, addr))
            return
    L.mfsc_host_free(ctypes.c_void_p(addr))


def host_buffer(nbytes, L=None):
    """uint8 array of nbytes for data that will be uploaded: page-locked when a device is there (pooled), else plain numpy.
    L: the bound library (tests pass the emulation build, whose mfsc_host_alloc is malloc)."""
    if nbytes < _PIN_MIN or nbytes > _PIN_MAX:
        return np.empty(nbytes, dtype=np.uint8)
    if L is None:
        from . import _lib
        try:
            L = _lib.load()
        except Exception:
            return np.empty(nbytes, dtype=np.uint8)
    addr = size = None
    with _PIN_LOCK:
        fit = [(b, a) for b, a in _PIN_FREE if nbytes <= b <= 2 * nbytes + (1 << 24)]
        if fit:
            size, addr = min(fit)
            _PIN_FREE.remove((size, addr))
    if addr is None:
        size = (nbytes + (1 << 24) - 1) >> 24 << 24          # 16 MB granules, so that batches of similar size share blocks
        p = ctypes.c_void_p()
        if L.mfsc_host_alloc(ctypes.byref(p), size) != 0 or not p.value:
            return np.empty(nbytes, dtype=np.uint8)          # no device (CPU-only run) or no page-locked memory left
        addr = p.value
    block = (ctypes.c_uint8 * size).from_address(addr)
    weakref.finalize(block, _pin_release, L, size, addr).atexit = False      # at exit the process gives everything back anyway
    return np.frombuffer(block, dtype=np.uint8, count=nbytes)


def _names_from_spans(data, spans):
    """Header texts data[b_i:e_i] -> list[str]: gathered into one newline-separated buffer with numpy, decoded and split once
    (a Python loop over 25 000 records costs as much as parsing their 250 MB of sequence)."""
    n = len(spans) // 2
    if n == 0:
        return []
    b, e = spans[0::2], spans[1::2]
    lens = e - b
    dst0 = np.concatenate([[0], np.cumsum(lens + 1)[:-1]])           # where name i starts in the output (one separator after each)
    total = int(lens.sum())
    out = np.full(total + n, 10, dtype=np.uint8)
    if total:
        within = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
        out[np.repeat(dst0, lens) + within] = np.asarray(data)[np.repeat(b, lens) + within]
    names = out.tobytes().decode("utf-8").split("\n")
    names.pop()
    if len(names) != n:          # a header text that itself holds a newline cannot occur (headers are single lines); be safe anyway
        raw = memoryview(data)
        sp = spans.tolist()
        names = [str(raw[sp[2 * i]:sp[2 * i + 1]], "utf-8") for i in range(n)]
    return names


def read_fasta_arrays(path):
    """Ingest for the aligner: -> (names list[str], uint8 bases, int64 offsets[n+1]).  Same records as read_fasta
    (blank lines skipped, sequence lines joined); one pass in C (mfsc_fasta_parse), no per-line Python work."""
    import ctypes
    from . import _lib
    L = _lib.load()
    size = os.path.getsize(path)
    data = np.memmap(path, dtype=np.uint8, mode="r") if size > (1 << 24) else np.fromfile(path, dtype=np.uint8)   # big files: parse straight from the page cache
    n_rec = ctypes.c_int32(0)
    _lib.check(L, L.mfsc_fasta_count(_lib.ptr(data), data.size, ctypes.byref(n_rec)))
    n = n_rec.value
    bases = host_buffer(max(int(data.size), 1))
    off = np.zeros(n + 1, dtype=np.int64)
    spans = np.zeros(2 * max(n, 1), dtype=np.int64)
    _lib.check(L, L.mfsc_fasta_parse(_lib.ptr(data), data.size, _lib.ptr(bases), _lib.ptr(off), _lib.ptr(spans), n, ctypes.byref(n_rec)))
    names = _names_from_spans(data, spans[:2 * n])
    return names, bases[:int(off[n])], off


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
    return part_boundaries_len([len(n) for n in names], [len(s) for s in seqs], n_parts, line_len)


def part_boundaries_len(name_lens, seq_lens, n_parts, line_len=60):
    """fasta-splitter.pl:160-183 on arrays.  The Perl loop opens a new part in front of record i (when the current part already
    holds a record) as soon as  total_i > part_end  and  total_i - part_end > part_end - total_{i-1},  total_i being the
    bytes written up to and including record i; with integer totals and part_end that is  total_i + total_{i-1} > 2 part_end,
    so every part boundary is one binary search over that (ascending) sum instead of a Python loop over all records."""
    name_lens = np.asarray(name_lens, dtype=np.int64)
    seq_lens = np.asarray(seq_lens, dtype=np.int64)
    n = len(seq_lens)
    if n == 0:
        return []
    lines = (seq_lens + line_len - 1) // line_len if line_len else np.ones(n, dtype=np.int64)
    sizes = seq_lens + name_lens + 1 + (1 + lines)
    tot = np.cumsum(sizes)
    total = int(tot[-1])
    pair = tot.copy()
    pair[1:] += tot[:-1]                     # total_i + total_{i-1}
    parts, begin, part = [], 0, 1
    while True:
        part_end = int(total / n_parts * part) + 1
        # first record i > begin with pair[i] > 2 * part_end
        i = int(np.searchsorted(pair, 2 * part_end, side="right"))
        if i <= begin:
            i = begin + 1
        if i >= n:
            parts.append((begin, n))
            return parts
        parts.append((begin, i))
        begin = i
        part += 1


def _part_boundaries_loop(name_lens, seq_lens, n_parts, line_len=60):
    """The Perl loop spelled out record by record (kept as the checker of part_boundaries_len in tests/)."""
    def seq_size(nlen, slen):
        return slen + nlen + 1 + (1 + ((slen + line_len - 1) // line_len if line_len else 1))
    sizes = [seq_size(int(n), int(s)) for n, s in zip(name_lens, seq_lens)]
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


def split_fasta(path, n_parts, out_dir=None, line_len=60, submit=None, keep=None):
    """Write <base>.part-NN.fasta files; returns their paths.  fasta-splitter.pl writes into
    the CWD and alignerRobot.py:268-269 then copies them into the folder; callers pass
    out_dir=folder to land them there directly.  Parsing and the wrapped part files are done in C
    (mfsc_fasta_parse / mfsc_fasta_write); the part boundaries follow the Perl script.
    submit(fn, *args): when given, the part files are written through it (a background pool) instead of inline;
    keep: dict that receives abspath(part) -> (names, bases, offsets) of every part, views into the parsed arrays."""
    from . import _lib
    L = _lib.load()
    names, bases, off = read_fasta_arrays(path)
    enc = [n.encode() for n in names]
    name_off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(e) for e in enc], out=name_off[1:])
    blob = b"".join(enc) + b"\0"
    base = os.path.basename(path)
    ext = ""
    for e in (".fasta", ".faa", ".fna", ".fa"):
        if base.lower().endswith(e) and len(base) > len(e):
            base, ext = base[:-len(e)], base[-len(e):]
            break
    out_dir = out_dir if out_dir is not None else os.getcwd()
    num_len = len(str(n_parts))
    paths = []
    bounds = part_boundaries_len(np.diff(name_off), np.diff(off), n_parts, line_len)
    bases = np.ascontiguousarray(bases)
    for p, (b, e) in enumerate(bounds, start=1):
        stem = base if (".part-" in base and base.rsplit(".part-", 1)[1].isdigit()) else base + ".part"
        out = os.path.join(out_dir, "%s-%0*d%s" % (stem, num_len, p, ext))
        def write(out=out, b=b, e=e):
            _lib.check(L, L.mfsc_fasta_write(out.encode(), blob, _lib.ptr(name_off), _lib.ptr(bases) if bases.size else _lib.ptr(np.zeros(1, np.uint8)),
                                             _lib.ptr(off), b, e, line_len))
        if submit is not None:
            submit(write)
        else:
            write()
        if keep is not None:
            keep[os.path.abspath(out)] = (names[b:e], bases[off[b]:off[e]], off[b:e + 1] - off[b])
        paths.append(out)
    return paths
