"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Restates, on the CPU, what `nucmer` + `show-coords -r` compute for the reference's
aligner wrapper (srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py:46-72)
and the Python loops that consume the rows (merger.py:101-190, abunSplitter.py:260-351).
PARITY UNPINNED against real MUMmer 3.23 -- see mfsc_oracle.cpp.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmfsc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mfsc_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("minmatch", "mincluster", "maxgap", "breaklen", "diagdiff", "maxmatch", "simplify", "band_cap")] + \
               [("diagfactor", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_run.restype = ctypes.c_void_p
        L.orc_run.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p,
                              ctypes.c_int, ctypes.POINTER(_Params), ctypes.c_int]
        L.orc_free.argtypes = [ctypes.c_void_p]
        for f in ("orc_n_mems", "orc_n_clusters", "orc_n_cluster_matches", "orc_n_rows", "orc_n_deltas"):
            getattr(L, f).restype = ctypes.c_int64
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.orc_get_mems.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_clusters.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_ref_build.restype = ctypes.c_void_p
        L.orc_ref_build.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_ref_free.argtypes = [ctypes.c_void_p]
        L.orc_run_ix.restype = ctypes.c_void_p
        L.orc_run_ix.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_Params), ctypes.c_int]
        L.orc_dp.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p]
        L.orc_coverage.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def make_params(minmatch=20, mincluster=65, maxgap=90, breaklen=200, diagdiff=5, diagfactor=0.12,
                maxmatch=True, simplify=True, band_cap=1024):
    return _Params(minmatch, mincluster, maxgap, breaklen, diagdiff, int(maxmatch), int(simplify), band_cap, diagfactor)


def _concat(seqs):
    """list of bytes/str -> (bytes, int64 offsets)"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bs], out=off[1:])
    return b"".join(bs), off


class Ref:
    """Reference records loaded and indexed once; `align(ref, reads)` calls may then share it from many threads
    (ctypes releases the GIL), which is how bench.py's CPU arm uses all host cores without rebuilding the index."""

    def __init__(self, ref_seqs, minmatch=20):
        rb, roff = _concat(ref_seqs)
        self.n = len(ref_seqs)
        self.h = lib().orc_ref_build(rb, roff.ctypes.data, self.n, minmatch)

    def free(self):
        if self.h:
            lib().orc_ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def dp(a, b, breaklen=200, band_cap=1 << 20, forced=False, optimal=False):
    """One run of the oracle's DP engine on two sequences (forward from their first bases).  Returns
    dict(reached, fi, fj, cells, max_band, ops) -- ops: 'M' pair column, 'I' a-only, 'D' b-only."""
    a = a.encode() if isinstance(a, str) else bytes(a)
    b = b.encode() if isinstance(b, str) else bytes(b)
    out = np.zeros(8, dtype=np.int64)
    buf = ctypes.create_string_buffer(len(a) + len(b) + 8)
    lib().orc_dp(a, len(a), b, len(b), breaklen, band_cap, int(forced), int(optimal), out.ctypes.data, buf)
    return dict(reached=bool(out[0]), fi=int(out[1]), fj=int(out[2]), cells=int(out[3]), max_band=int(out[4]), ops=buf.value.decode())


def align(ref_seqs, qry_seqs, stage_limit=3, **kw):
    """Run the oracle.  Returns dict with
    mems     int64 [n,5]  q, strand, s1 (global ref pos), s2, len     sorted (q, strand, s2, s1)
    clusters (cl int64 [n,4] q, strand, first, count ; cm int64 [m,3] s1, s2, len)
    rows     int64 [n,11] ref, qry, dir, sA, eA, sB, eB, errors, cols, delta_begin, delta_end (delta-file order)
    deltas   int64 [k]
    stats    dict(cells, max_band, dp_calls)
    """
    L = lib()
    P = make_params(**kw)
    qb, qoff = _concat(qry_seqs)
    if isinstance(ref_seqs, Ref):
        h = L.orc_run_ix(ref_seqs.h, qb, qoff.ctypes.data, len(qry_seqs), ctypes.byref(P), stage_limit)
        if not h:
            raise ValueError("the shared reference index was built for a larger minmatch")
    else:
        rb, roff = _concat(ref_seqs)
        h = L.orc_run(rb, roff.ctypes.data, len(ref_seqs), qb, qoff.ctypes.data, len(qry_seqs), ctypes.byref(P), stage_limit)
    try:
        out = {}
        n = L.orc_n_mems(h)
        mems = np.zeros((n, 5), dtype=np.int64)
        L.orc_get_mems(h, mems.ctypes.data)
        out["mems"] = mems
        nc, ncm = L.orc_n_clusters(h), L.orc_n_cluster_matches(h)
        cl = np.zeros((nc, 4), dtype=np.int64)
        cm = np.zeros((ncm, 3), dtype=np.int64)
        L.orc_get_clusters(h, cl.ctypes.data, cm.ctypes.data)
        out["clusters"] = (cl, cm)
        nr, nd = L.orc_n_rows(h), L.orc_n_deltas(h)
        rows = np.zeros((nr, 11), dtype=np.int64)
        deltas = np.zeros(nd, dtype=np.int64)
        L.orc_get_rows(h, rows.ctypes.data, deltas.ctypes.data)
        out["rows"] = rows
        out["deltas"] = deltas
        st = np.zeros(4, dtype=np.int64)
        L.orc_get_stats(h, st.ctypes.data)
        out["stats"] = dict(cells=int(st[0]), max_band=int(st[1]), dp_calls=int(st[2]), cap_hits=int(st[3]))
        return out
    finally:
        L.orc_free(h)


def idy_of(errors, cols):
    """show-coords arithmetic: single precision (total - errors) / total * 100."""
    t = np.float32(cols)
    return np.float32((t - np.float32(errors)) / t * np.float32(100.0))


def coords_rows(rows, ref_names, qry_names):
    """delta-order oracle rows -> show-coords -r ordered list of
    [S1, E1, S2, E2, LEN1, LEN2, IDY(2dp float), refName, qryName]  (alignerRobot.py:34)."""
    order = sorted(range(len(rows)), key=lambda i: (ref_names[rows[i][0]], rows[i][3]))  # stable: ties keep delta order
    out = []
    for i in order:
        r = rows[i]
        idy = float("%.2f" % float(idy_of(r[7], r[8])))
        out.append([int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(abs(r[4] - r[3]) + 1), int(abs(r[6] - r[5]) + 1),
                    idy, ref_names[r[0]], qry_names[r[1]]])
    return out


def coords_text(rows, ref_names, qry_names, ref_path, qry_path):
    """The text `show-coords -r` would print (5 header lines, alignerRobot.py:12-13;
    row layout toCondenseFixer.py:133-139)."""
    lines = ["%s %s" % (ref_path, qry_path), "NUCMER", "",
             "    [S1]     [E1]  |     [S2]     [E2]  |  [LEN 1]  [LEN 2]  |  [% IDY]  | [TAGS]",
             "=" * 85]
    for r in coords_rows(rows, ref_names, qry_names):
        lines.append("%8d %8d  | %8d %8d  | %8d %8d  | %8.2f  | %s\t%s" % tuple(r))
    return "\n".join(lines) + "\n"


def coverage(ref_id, s1, e1, contig_off):
    """merger.py:126-140 restated: +1 over [S1-1, E1) per selected row."""
    L = lib()
    ref_id = np.ascontiguousarray(ref_id, dtype=np.int32)
    s1 = np.ascontiguousarray(s1, dtype=np.int32)
    e1 = np.ascontiguousarray(e1, dtype=np.int32)
    contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
    cov = np.zeros(int(contig_off[-1]), dtype=np.int32)
    L.orc_coverage(ref_id.ctypes.data, s1.ctypes.data, e1.ctypes.data, len(s1), contig_off.ctypes.data,
                   len(contig_off) - 1, cov.ctypes.data)
    return cov
