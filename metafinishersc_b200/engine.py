"""Python handles over the C ABI: contig index, read batches, row sets, coverage.

Thin by design -- numpy buffers in, numpy buffers out; all computation happens in
libmfsc_b200.so on the GPU (include/mfsc.h).
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import KEEP_CLUSTERS, KEEP_MEMS, STOP_CLUSTERS, STOP_SEEDS, WANT_DELTAS, check, make_params, ptr  # noqa: F401


def _as_buf(seqs):
    """list[bytes] or (uint8 array, int64 offsets) -> (uint8 array, int64 offsets)"""
    if isinstance(seqs, tuple):
        buf, off = seqs
        return np.ascontiguousarray(buf, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.int64)
    return _lib.concat([s.encode() if isinstance(s, str) else bytes(s) for s in seqs])


class Rows:
    """Alignment rows of one batch in delta order, structure of arrays (int32)."""
    __slots__ = ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols", "stats", "ms", "mems", "clusters",
                 "delta_begin", "delta_end", "deltas")

    def __len__(self):
        return len(self.s1)

    def idy(self):
        """%IDY as show-coords computes it: single precision (cols - errors) / cols * 100."""
        t = self.cols.astype(np.float32)
        return (t - self.errors.astype(np.float32)) / t * np.float32(100.0)


class Index:
    def __init__(self, L, handle, n_rec):
        self.L, self.h, self.n_rec = L, handle, n_rec

    def info(self):
        a = np.zeros(8, dtype=np.int64)
        check(self.L, self.L.mfsc_index_info(self.h, ptr(a)))
        return dict(k=int(a[0]), device=int(a[1]), text_words=int(a[2]), occupied_buckets=int(a[3]), positions=int(a[4]), hbm_bytes=int(a[5]),
                    records=int(a[6]), bases=int(a[7]))

    def buffers(self):
        """[(device pointer, bytes)] x 5: packed text, invalid mask, bucket directory, positions, heads of the occupied buckets."""
        p = (ctypes.c_void_p * 5)()
        b = np.zeros(5, dtype=np.int64)
        check(self.L, self.L.mfsc_index_buffers(self.h, p, ptr(b)))
        return [(int(p[i] or 0), int(b[i])) for i in range(5)]

    def build_ms(self, tables_only=False):
        v = (ctypes.c_double * 2)()
        check(self.L, self.L.mfsc_index_build_ms(self.h, v))
        return v[1] if tables_only else v[0]

    def free(self):
        if self.h:
            self.L.mfsc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Reads:
    def __init__(self, L, handle, n_reads, n_bases):
        self.L, self.h, self.n_reads, self.n_bases = L, handle, n_reads, n_bases

    def free(self):
        if self.h:
            self.L.mfsc_reads_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CoverageProfile:
    """Per-base coverage of a contig set resident in HBM with 64-bit tile prefix sums (mfsc_cov_*): the consumer of
    merger.assignCoverage's lists, mergerHelper.formatIntervalAndFilter (mergerHelper.py:137-163), only sums intervals."""

    def __init__(self, L, h, contig_off, ms):
        self.L, self.h, self.off, self.ms = L, h, contig_off, ms

    def contig_len(self, c):
        return int(self.off[c + 1] - self.off[c])

    def sums(self, contig, start, end):
        """sum(cov[contig[q]][start[q]:end[q]]) for every q, Python slice semantics, int64."""
        contig = np.ascontiguousarray(contig, dtype=np.int32)
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        out = np.zeros(len(contig), dtype=np.int64)
        check(self.L, self.L.mfsc_cov_interval_sums(self.h, ptr(contig), ptr(start), ptr(end), len(contig), ptr(out)))
        return out

    def array(self, contig=-1):
        n = int(self.off[-1]) if contig < 0 else self.contig_len(contig)
        out = np.zeros(n, dtype=np.int32)
        check(self.L, self.L.mfsc_cov_fetch(self.h, contig, ptr(out)))
        return out

    def free(self):
        if self.h:
            self.L.mfsc_cov_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, lib=None):
        self.L = lib if lib is not None else _lib.load()

    def device_count(self):
        n = ctypes.c_int32()
        check(self.L, self.L.mfsc_device_count(ctypes.byref(n)))
        return n.value

    def set_device(self, d):
        check(self.L, self.L.mfsc_set_device(d))

    def trim_pool(self):
        """Gives the scratch kept in the device's memory pool back to the driver (mfsc_trim_pool)."""
        check(self.L, self.L.mfsc_trim_pool())

    def thread_release(self):
        """Destroys the calling thread's stream (worker threads call it before they end)."""
        self.L.mfsc_thread_release()

    # -- stage 1 ---------------------------------------------------------------------------
    def index(self, contigs, params=None):
        params = params or make_params()
        buf, off = _as_buf(contigs)
        h = ctypes.c_void_p()
        check(self.L, self.L.mfsc_index_build(ptr(buf), ptr(off), len(off) - 1, ctypes.byref(params), ctypes.byref(h)))
        return Index(self.L, h, len(off) - 1)

    # -- stages 2-4 ------------------------------------------------------------------------
    def upload(self, reads):
        buf, off = _as_buf(reads)
        h = ctypes.c_void_p()
        check(self.L, self.L.mfsc_reads_upload(ptr(buf), ptr(off), len(off) - 1, ctypes.byref(h)))
        return Reads(self.L, h, len(off) - 1, int(off[-1] - off[0]))

    def _collect(self, h, flags):
        L = self.L
        try:
            r = Rows()
            n = L.mfsc_result_count(h, 0)
            cols = [np.zeros(n, dtype=np.int32) for _ in range(8)]
            check(L, L.mfsc_result_rows(h, *[ptr(c) for c in cols]))
            r.ref_id, r.qry_id, r.s1, r.e1, r.s2, r.e2, r.errors, r.cols = cols
            st = np.zeros(16, dtype=np.int64)
            ms = np.zeros(16, dtype=np.float64)
            check(L, L.mfsc_result_stats(h, ptr(st), ptr(ms)))
            r.stats = dict(cells=int(st[0]), dp_calls=int(st[1]), max_band=int(st[2]), probes=int(st[3]), mems=int(st[4]),
                           launches=int(st[5]), seed_bytes=int(st[6]), h2d_bytes=int(st[7]), d2h_bytes=int(st[8]), dp_retry_reads=int(st[9]), cap_hits=int(st[10]),
                           gap_tasks=int(st[11]), gap_hits=int(st[12]))
            r.ms = dict(upload=ms[0], seeds=ms[1], sort=ms[2], cluster=ms[3], extend=ms[4], download=ms[5], total=ms[6])
            r.mems = r.clusters = None
            r.delta_begin = r.delta_end = r.deltas = None
            if flags & WANT_DELTAS:
                r.delta_begin, r.delta_end = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
                r.deltas = np.zeros(L.mfsc_result_count(h, 4), dtype=np.int32)
                check(L, L.mfsc_result_deltas(h, ptr(r.delta_begin), ptr(r.delta_end), ptr(r.deltas)))
            if flags & KEEP_MEMS:
                m = L.mfsc_result_count(h, 1)
                q, st_, s2, ln = [np.zeros(m, dtype=np.int32) for _ in range(4)]
                s1 = np.zeros(m, dtype=np.int64)
                check(L, L.mfsc_result_mems(h, ptr(q), ptr(st_), ptr(s1), ptr(s2), ptr(ln)))
                r.mems = np.stack([q.astype(np.int64), st_.astype(np.int64), s1, s2.astype(np.int64), ln.astype(np.int64)], axis=1)
            if flags & KEEP_CLUSTERS:
                nc, ncm = L.mfsc_result_count(h, 2), L.mfsc_result_count(h, 3)
                q, st_, rec, first, cnt = [np.zeros(nc, dtype=np.int32) for _ in range(5)]
                ms1 = np.zeros(ncm, dtype=np.int64)
                ms2, mln = np.zeros(ncm, dtype=np.int32), np.zeros(ncm, dtype=np.int32)
                check(L, L.mfsc_result_clusters(h, ptr(q), ptr(st_), ptr(rec), ptr(first), ptr(cnt), ptr(ms1), ptr(ms2), ptr(mln)))
                r.clusters = (np.stack([q, st_, rec, first, cnt], axis=1).astype(np.int64),
                              np.stack([ms1, ms2.astype(np.int64), mln.astype(np.int64)], axis=1))
            return r
        finally:
            L.mfsc_result_free(h)

    def align_reads(self, index, reads, params=None, flags=0):
        """Stages 2-4 on a batch already resident in HBM."""
        params = params or make_params()
        h = ctypes.c_void_p()
        check(self.L, self.L.mfsc_align_reads(index.h, reads.h, ctypes.byref(params), flags, ctypes.byref(h)))
        return self._collect(h, flags)

    # Engine.align pipelines a batch of at least two pieces of PIECE_BASES when the copy is a large share of the call.  Every piece
    # ends with the tail of its slowest reads (one warp per read), so pieces pay only where reads are cheap: measured, accurate 6 kbp
    # reads (27 Gbp/s on the device) 21.6 -> 20.1 ms in four pieces, 10 kbp reads at 15 % error (2 Gbp/s) 130 -> 137 ms in two.
    # The rate of the previous call on the same index decides; a first call goes up whole.
    PIECE_BASES = 96_000_000
    PIECE_MIN_RATE = 8e9          # bases per second of device time

    def align(self, index, reads, params=None, flags=0, pieces=None):
        """Host buffers in, rows out: what one `nucmer` + `show-coords` job produced.  A large batch goes up in pieces on a
        second host thread / stream, so that the upload (and packing) of piece k+1 overlaps the alignment of piece k; every
        read is aligned on its own, so the rows are those of one call.  pieces: None = by size and the device rate of the previous
        call on this index, 1 = one call."""
        params = params or make_params()
        buf, off = _as_buf(reads)
        if len(off) <= 1:
            r = Rows()
            z = np.zeros(0, dtype=np.int32)
            r.ref_id = r.qry_id = r.s1 = r.e1 = r.s2 = r.e2 = r.errors = r.cols = z
            r.stats, r.ms, r.mems, r.clusters = {}, {}, None, None
            r.delta_begin = r.delta_end = r.deltas = None
            return r
        n = len(off) - 1
        if pieces is None and os.environ.get("MFSC_ALIGN_PIECES"):
            pieces = int(os.environ["MFSC_ALIGN_PIECES"])         # experiments
        if pieces is None:
            bases = int(off[-1] - off[0])
            pieces = min(4, bases // self.PIECE_BASES) if getattr(index, "_rate", 0.0) >= self.PIECE_MIN_RATE else 1
        if flags & (KEEP_MEMS | KEEP_CLUSTERS | STOP_SEEDS | STOP_CLUSTERS):
            pieces = 1
        pieces = max(1, min(pieces, n))
        if pieces <= 1:
            h = ctypes.c_void_p()
            check(self.L, self.L.mfsc_align_batch(index.h, ptr(buf), ptr(off), n, ctypes.byref(params), flags, ctypes.byref(h)))
            r = self._collect(h, flags)
        else:
            r = self._align_pipelined(index, buf, off, params, flags, pieces)
        if flags == 0 or flags == WANT_DELTAS:
            dev_s = (r.ms.get("total", 0.0) - r.ms.get("upload", 0.0)) * 1e-3
            if dev_s > 0:
                index._rate = float(off[-1] - off[0]) / dev_s
        return r

    PIECE_WORKERS = 2             # host threads (streams) that align pieces: the tail of one piece overlaps the start of the next

    def _align_pipelined(self, index, buf, off, params, flags, pieces):
        import queue
        import threading
        n = len(off) - 1
        # cut by bases, at read boundaries
        targets = off[0] + (off[-1] - off[0]) * np.arange(1, pieces) // pieces
        cuts = [0] + [int(x) for x in np.searchsorted(off, targets)] + [n]
        cuts = sorted(set(cuts))
        n_pieces = len(cuts) - 1
        workers = max(1, min(self.PIECE_WORKERS, n_pieces))
        q = queue.Queue()
        dev = index.info()["device"]
        parts = [None] * n_pieces
        errors = []

        def up():
            E2 = None
            try:
                E2 = Engine(self.L)
                E2.set_device(dev)
                for k in range(n_pieces):
                    lo, hi = cuts[k], cuts[k + 1]
                    q.put((k, E2.upload((buf[off[lo] - off[0]:off[hi] - off[0]], off[lo:hi + 1] - off[lo]))))
            except Exception as e:        # surfaced in the caller's thread
                errors.append(e)
            finally:
                for _ in range(workers):
                    q.put(None)
                if E2 is not None:
                    E2.thread_release()

        def work(E, own_thread):
            try:
                if own_thread:
                    E.set_device(dev)
                while True:
                    item = q.get()
                    if item is None:
                        break
                    k, rd = item
                    try:
                        if not errors:
                            parts[k] = E.align_reads(index, rd, params, flags)
                    except Exception as e:
                        errors.append(e)
                    finally:
                        rd.free()
            finally:
                if own_thread:
                    E.thread_release()

        threads = [threading.Thread(target=up)] + [threading.Thread(target=work, args=(Engine(self.L), True)) for _ in range(workers - 1)]
        for th in threads:
            th.start()
        try:
            work(self, False)
        finally:
            for th in threads:
                th.join()
        if errors:
            raise errors[0]
        r = Rows()
        for f in ("ref_id", "s1", "e1", "s2", "e2", "errors", "cols"):
            setattr(r, f, np.concatenate([getattr(p, f) for p in parts]))
        r.qry_id = np.concatenate([p.qry_id + cuts[k] for k, p in enumerate(parts)])
        r.stats = {}
        for key in parts[0].stats:
            vals = [p.stats[key] for p in parts]
            r.stats[key] = max(vals) if key == "max_band" else sum(vals)
        r.stats["h2d_bytes"] = int(off[-1] - off[0]) + 8 * (n + 1) + 28 * n
        r.stats["launches"] += len(parts)
        r.ms = {key: sum(p.ms[key] for p in parts) for key in parts[0].ms}
        r.mems = r.clusters = None
        r.delta_begin = r.delta_end = r.deltas = None
        if flags & WANT_DELTAS:
            base, db, de, dl = 0, [], [], []
            for p in parts:
                ok = p.delta_begin >= 0
                db.append(np.where(ok, p.delta_begin + base, -1)); de.append(np.where(ok, p.delta_end + base, -1)); dl.append(p.deltas)
                base += len(p.deltas)
            r.delta_begin, r.delta_end, r.deltas = np.concatenate(db), np.concatenate(de), np.concatenate(dl)
        return r

    # -- stage 5 ---------------------------------------------------------------------------
    def coverage(self, ref_id, s1, e1, contig_off, want_ms=False):
        ref_id = np.ascontiguousarray(ref_id, dtype=np.int32)
        s1 = np.ascontiguousarray(s1, dtype=np.int32)
        e1 = np.ascontiguousarray(e1, dtype=np.int32)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
        cov = np.zeros(int(contig_off[-1]), dtype=np.int32)
        ms = np.zeros(2, dtype=np.float64)
        check(self.L, self.L.mfsc_coverage(ptr(ref_id), ptr(s1), ptr(e1), len(s1), ptr(contig_off), len(contig_off) - 1,
                                           ptr(cov), ptr(ms)))
        return (cov, ms) if want_ms else cov

    def coverage_profile(self, ref_id, s1, e1, contig_off):
        """Stage 5 kept resident in HBM: a CoverageProfile answering interval sums on the device."""
        ref_id = np.ascontiguousarray(ref_id, dtype=np.int32)
        s1 = np.ascontiguousarray(s1, dtype=np.int32)
        e1 = np.ascontiguousarray(e1, dtype=np.int32)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
        h = ctypes.c_void_p()
        ms = np.zeros(2, dtype=np.float64)
        check(self.L, self.L.mfsc_cov_build(ptr(ref_id), ptr(s1), ptr(e1), len(s1), ptr(contig_off), len(contig_off) - 1,
                                            ctypes.byref(h), ptr(ms)))
        return CoverageProfile(self.L, h, contig_off, ms)

    def row_flags(self, rows, ref_len, qry_len, ref_gid, qry_gid):
        """MFSC_RF_* predicate bits of every row (mfsc_row_flags)."""
        rl, ql = np.ascontiguousarray(ref_len, dtype=np.int32), np.ascontiguousarray(qry_len, dtype=np.int32)
        rg, qg = np.ascontiguousarray(ref_gid, dtype=np.int32), np.ascontiguousarray(qry_gid, dtype=np.int32)
        out = np.zeros(len(rows), dtype=np.uint32)
        check(self.L, self.L.mfsc_row_flags(len(rows), ptr(rows.ref_id), ptr(rows.qry_id), ptr(rows.s1), ptr(rows.e1), ptr(rows.s2),
                                            ptr(rows.e2), ptr(rl), len(rl), ptr(ql), len(ql), ptr(rg), ptr(qg), ptr(out)))
        return out

    # -- text outputs ----------------------------------------------------------------------
    def write_coords(self, path, rows, ref_names, qry_names, ref_path, qry_path, sort_by_ref=True):
        rn = b"\0".join(n.encode() for n in ref_names) + b"\0"
        qn = b"\0".join(n.encode() for n in qry_names) + b"\0"
        check(self.L, self.L.mfsc_write_coords(path.encode(), ref_path.encode(), qry_path.encode(), len(rows), ptr(rows.ref_id),
                                               ptr(rows.qry_id), ptr(rows.s1), ptr(rows.e1), ptr(rows.s2), ptr(rows.e2),
                                               ptr(rows.errors), ptr(rows.cols), rn, len(ref_names), qn, len(qry_names),
                                               int(sort_by_ref)))

    def write_delta(self, path, rows, ref_names, ref_len, qry_names, qry_len, ref_path, qry_path):
        rn = b"\0".join(n.encode() for n in ref_names) + b"\0"
        qn = b"\0".join(n.encode() for n in qry_names) + b"\0"
        rl = np.ascontiguousarray(ref_len, dtype=np.int32)
        ql = np.ascontiguousarray(qry_len, dtype=np.int32)
        check(self.L, self.L.mfsc_write_delta(path.encode(), ref_path.encode(), qry_path.encode(), len(rows), ptr(rows.ref_id),
                                              ptr(rows.qry_id), ptr(rows.s1), ptr(rows.e1), ptr(rows.s2), ptr(rows.e2),
                                              ptr(rows.errors), ptr(rows.cols), rn, ptr(rl), len(ref_names), qn, ptr(ql), len(qry_names),
                                              ptr(getattr(rows, "delta_begin", None)), ptr(getattr(rows, "delta_end", None)),
                                              ptr(getattr(rows, "deltas", None))))
