"""Drop-in for the reference's aligner wrapper, same function names and argument meaning
(reference srcRefactor/repeatPhaserLib/finisherSCCoreLib/alignerRobot.py):

    extractMumData :7     useMummerAlign :42    nucmerMummer :46     showCoorMummer :66
    combineMultipleCoorMum :74   useMummerAlignBatch :114   transformCoor :242   largeRvsQAlign :258

Where the reference runs `os.system("nucmer ...")` / `os.system("show-coords -r ...")`, these functions
read the same FASTA paths, run stages 1-4 on the GPU through libmfsc_b200.so and write the same files
(`<folder><out>.delta`, `<folder><out>Out` or `<folder><specialName>`), so every consumer that opens
a coords file keeps working.  `mummerLink` is accepted and ignored.  The contig index is cached
across calls, so the 20 read parts of one reads-vs-contigs alignment share one index instead of
rebuilding it 20 times.  There is no CPU path: without the CUDA library every aligning call raises.
"""
import hashlib
import os
import threading

import numpy as np

from . import engine as _engine_mod
from . import fasta, houseKeeper

_ENGINE = None
_INDEX_CACHE = {}      # (content signature of the reference file, minmatch) -> (Index, names, lens)
_ROWS_CACHE = {}       # delta path -> (RowSet, stat signature of the delta file this process wrote)
_CACHE_LOCK = threading.RLock()     # the GPU fan-out (_batch_on_gpus) touches the caches from several host threads
_INDEX_CACHE_MAX = 2
_ROWS_CACHE_MAX = 64
_HASH_WHOLE_BELOW = 64 << 20        # reference files up to this size are identified by a hash of all their bytes
_QUERY_CACHE = {}      # query part path -> (names, bases, offsets) handed from largeRvsQAlign's splitter to the jobs in memory
_WRITERS = None        # background threads for the text files (part FASTA, .delta, coords): the C writers release the GIL
_PENDING = []          # futures of file writes not yet joined; every public entry point joins them before it returns


def _submit(fn, *args):
    """Run a file write on the writer pool (overlaps with the alignment of the next part)."""
    global _WRITERS
    if _WRITERS is None:
        from concurrent.futures import ThreadPoolExecutor
        _WRITERS = ThreadPoolExecutor(max_workers=4)
    f = _WRITERS.submit(fn, *args)
    with _CACHE_LOCK:
        _PENDING.append(f)
    return f


def _join_writes():
    """All files of the calls so far are on disk when this returns (errors of the writers surface here)."""
    while True:
        with _CACHE_LOCK:
            if not _PENDING:
                return
            f = _PENDING.pop(0)
        f.result()
fullDelta = False      # True: <prefix>.delta holds real MUMmer delta records (indel offset lists; the traceback-recording kernel
                       # and ~0.6 B of text per query base, about 3x slower end to end).  False (default): rows-only records under a
                       # `#MFSC rows-only` marker line -- nothing in the reference parses .delta files, show-coords is
                       # served from the rows


def set_engine(E):
    """Bind the wrapper to an engine (tests bind the host-emulation build; default is the CUDA library)."""
    global _ENGINE
    _ENGINE = E
    clear_caches()


def get_engine():
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = _engine_mod.Engine()
    return _ENGINE


def clear_caches():
    with _CACHE_LOCK:
        for ix, _, _ in _INDEX_CACHE.values():
            ix.free()
        _INDEX_CACHE.clear()
        _ROWS_CACHE.clear()


_BATCH_SIGS = None     # inside one batch call: path -> signature (a reference shared by 20 jobs is hashed once)


def _content_signature(path):
    memo = _BATCH_SIGS
    if memo is not None:
        key = os.path.abspath(path)
        if key not in memo:
            memo[key] = _content_signature_now(path)
        return memo[key]
    return _content_signature_now(path)


def _content_signature_now(path):
    """Identifies the CONTENT of a reference file, not its path or time stamp: IORobot.alignWithName rewrites
    `<name>leftSeg.fasta` in place on every call, usually with the same size, and a coarse mtime (1 s, NFS) would
    make (path, mtime, size) collide.  Files up to 64 MB are hashed whole (a few ms); larger ones by size, inode,
    ctime/mtime and a hash of their first and last MB."""
    st = os.stat(path)
    h = hashlib.blake2b(digest_size=16)
    with open(path, "rb") as f:
        if st.st_size <= _HASH_WHOLE_BELOW:
            h.update(f.read())
            return ("all", st.st_size, h.hexdigest())
        h.update(f.read(1 << 20))
        f.seek(st.st_size - (1 << 20))
        h.update(f.read(1 << 20))
    return ("ends", st.st_size, st.st_ino, st.st_ctime_ns, st.st_mtime_ns, h.hexdigest())


def _stat_signature(path):
    st = os.stat(path)
    return (st.st_size, st.st_ino, st.st_mtime_ns, st.st_ctime_ns)


def _rows_cache_put(delta, rs):
    _PENDING_ROWS.pop(delta, None)
    with _CACHE_LOCK:
        while len(_ROWS_CACHE) >= _ROWS_CACHE_MAX and delta not in _ROWS_CACHE:
            _ROWS_CACHE.pop(next(iter(_ROWS_CACHE)), None)
        _ROWS_CACHE[delta] = (rs, _stat_signature(delta) if os.path.exists(delta) else None)


def _rows_cache_get(delta):
    """Rows this process produced for `delta`, as long as the file on disk is still the one it wrote."""
    with _CACHE_LOCK:
        hit = _ROWS_CACHE.get(delta)
    if hit is None:
        return None
    rs, sig = hit
    try:
        if sig is not None and _stat_signature(delta) == sig:
            return rs
    except OSError:
        pass
    with _CACHE_LOCK:
        _ROWS_CACHE.pop(delta, None)
    return None


class RowSet:
    """Rows of one nucmer job plus what show-coords needs to print them."""

    def __init__(self, rows, ref_names, ref_lens, qry_names, qry_lens, ref_path, qry_path):
        self.rows, self.ref_names, self.ref_lens = rows, ref_names, ref_lens
        self.qry_names, self.qry_lens, self.ref_path, self.qry_path = qry_names, qry_lens, ref_path, qry_path

    def order(self):
        """show-coords -r order: reference id (string compare), then S1; ties keep delta order."""
        r = self.rows
        if len(r) == 0:
            return np.zeros(0, dtype=np.int64)
        uniq = sorted(set(self.ref_names))
        rank = {n: i for i, n in enumerate(uniq)}
        rr = np.array([rank[n] for n in self.ref_names], dtype=np.int64)[r.ref_id]
        return np.lexsort((np.arange(len(r)), r.s1, rr))

    def datalist(self):
        """The list extractMumData would parse from the coords text (alignerRobot.py:34).  Built from the row arrays in
        bulk: %IDY is the single-precision show-coords value rounded to two decimals the way "%.2f" prints it (the float32
        times 100 is exact in double precision, so rint / 100 is the correctly rounded decimal)."""
        r = self.rows
        if len(r) == 0:
            return []
        o = self.order()
        s1, e1, s2, e2 = r.s1[o].astype(np.int64), r.e1[o].astype(np.int64), r.s2[o].astype(np.int64), r.e2[o].astype(np.int64)
        idy = (np.rint(r.idy()[o].astype(np.float64) * 100.0) / 100.0).tolist()
        rn, qn = self.ref_names, self.qry_names
        rnames = [rn[i] for i in r.ref_id[o].tolist()]
        qnames = [qn[i] for i in r.qry_id[o].tolist()]
        return [list(t) for t in zip(s1.tolist(), e1.tolist(), s2.tolist(), e2.tolist(), (np.abs(e1 - s1) + 1).tolist(),
                                     (np.abs(e2 - s2) + 1).tolist(), idy, rnames, qnames)]


def _params(refinedVersion, maxmatch=True):
    minmatch = 10 if refinedVersion else 20
    breaklen = 50 if (houseKeeper.globalFast and not refinedVersion) else 200
    return _engine_mod.make_params(minmatch=minmatch, breaklen=breaklen, maxmatch=maxmatch)


def _get_index(E, ref_path, params):
    key = (_content_signature(ref_path), params.minmatch)
    with _CACHE_LOCK:
        hit = _INDEX_CACHE.get(key)
        if hit is not None:
            return hit
        names, buf, off = fasta.read_fasta_arrays(ref_path)
        names = [n.split()[0] if n.split() else n for n in names]      # nucmer ids stop at the first blank
        while len(_INDEX_CACHE) >= _INDEX_CACHE_MAX:
            old = next(iter(_INDEX_CACHE))
            _INDEX_CACHE.pop(old)[0].free()
        ix = E.index((buf, off), params)
        _INDEX_CACHE[key] = (ix, names, np.diff(off).tolist())
        return _INDEX_CACHE[key]


def _run_nucmer(prefix, ref_path, qry_path, params, E=None, index_entry=None):
    """One nucmer job: rows cached under `<prefix>.delta` and the delta file written (on the writer pool)."""
    E = E or get_engine()
    delta = prefix + ".delta"
    with _CACHE_LOCK:
        parsed = _QUERY_CACHE.get(os.path.abspath(qry_path))
    if not (os.path.exists(ref_path) and (parsed is not None or os.path.exists(qry_path))):
        print("ERROR: Could not find input file(s) %s %s" % (ref_path, qry_path))   # nucmer's failure is ignored upstream too
        with _CACHE_LOCK:
            _ROWS_CACHE.pop(delta, None)
        if os.path.exists(delta):
            os.remove(delta)
        return None
    ix, rnames, rlens = index_entry if index_entry is not None else _get_index(E, ref_path, params)
    qnames, qbuf, qoff = parsed if parsed is not None else fasta.read_fasta_arrays(qry_path)
    qnames = [n.split()[0] if n.split() else n for n in qnames]
    try:
        rows = E.align(ix, (qbuf, qoff), params, flags=_engine_mod.WANT_DELTAS if fullDelta else 0)
    except _engine_mod._lib.MfscError as err:
        if not fullDelta or "delta lists" not in str(err):
            raise
        print("note: delta lists did not fit for %s, writing alignment headers only (%s)" % (qry_path, err))
        rows = E.align(ix, (qbuf, qoff), params)
    rs = RowSet(rows, rnames, rlens, qnames, np.diff(qoff).tolist(), ref_path, qry_path)
    with _CACHE_LOCK:
        _ROWS_CACHE.pop(delta, None)

    def write():
        E.write_delta(delta, rows, rnames, rlens, qnames, rs.qry_lens, os.path.abspath(ref_path), os.path.abspath(qry_path))
        _rows_cache_put(delta, rs)
    _PENDING_ROWS[delta] = rs          # served from here until the file is on disk
    _submit(write)
    return rs


_PENDING_ROWS = {}     # delta path -> RowSet whose delta file is still being written


def nucmerMummer(specialForRaw, mummerLink, folderName, outputName, referenceName, queryName, refinedVersion):
    """alignerRobot.py:46-64: `nucmer [-b 50 | -l 10] --maxmatch -p <folder><out> <folder><ref> <folder|''><qry>`."""
    qry = queryName if specialForRaw else folderName + queryName
    rs = _run_nucmer(folderName + outputName, folderName + referenceName, qry, _params(refinedVersion))
    _join_writes()
    return rs


def nucmerDefault(mummerLink, prefix, ref_path, qry_path):
    """The raw default-mode call of adaptorFix.py:21 (`nucmer -p <prefix> <ref> <qry>`, no --maxmatch)."""
    rs = _run_nucmer(prefix, ref_path, qry_path, _params(False, maxmatch=False))
    _join_writes()
    return rs


def _load_delta(delta):
    """Rows of a delta file this module wrote earlier (resume paths); None when the file is absent.
    Reads both forms mfsc_write_delta produces: full records (7th header field 0, indel offsets, terminating 0)
    and header-only records (7th field = alignment columns)."""
    if not os.path.exists(delta):
        return None
    with open(delta) as f:
        head = f.readline().split()
        f.readline()
        rnames, qnames, rlens, qlens, rid, qid = [], [], [], [], {}, {}
        cols = [[] for _ in range(8)]
        cur = None
        for line in f:
            if line.startswith("#"):
                continue
            if line.startswith(">"):
                a = line[1:].split()
                if a[0] not in rid:
                    rid[a[0]] = len(rnames); rnames.append(a[0]); rlens.append(int(a[2]))
                if a[1] not in qid:
                    qid[a[1]] = len(qnames); qnames.append(a[1]); qlens.append(int(a[3]))
                cur = (rid[a[0]], qid[a[1]])
            else:
                a = line.split()
                if len(a) == 7:
                    s1, e1 = int(a[0]), int(a[1])
                    ncol = int(a[6]) if int(a[6]) > 0 else abs(e1 - s1) + 1       # full record: + one column per read-only base below
                    for c, v in zip(cols, (cur[0], cur[1], s1, e1, int(a[2]), int(a[3]), int(a[4]), ncol)):
                        c.append(v)
                elif len(a) == 1 and int(a[0]) < 0:
                    cols[7][-1] += 1
    rows = _engine_mod.Rows()
    (rows.ref_id, rows.qry_id, rows.s1, rows.e1, rows.s2, rows.e2, rows.errors, rows.cols) = \
        [np.array(c, dtype=np.int32) for c in cols]
    rows.stats, rows.ms, rows.mems, rows.clusters = {}, {}, None, None
    rows.delta_begin = rows.delta_end = rows.deltas = None
    return RowSet(rows, rnames, rlens, qnames, qlens, head[0] if head else "", head[1] if len(head) > 1 else "")


def _rowset_for(delta):
    rs = _PENDING_ROWS.get(delta)
    if rs is not None:
        return rs
    rs = _rows_cache_get(delta)
    return rs if rs is not None else _load_delta(delta)


def _write_coords(rs, out_path):
    if rs is None:
        open(out_path, "w").close()     # `show-coords missing.delta > out` leaves an empty file
        return
    get_engine().write_coords(out_path, rs.rows, rs.ref_names, rs.qry_names, rs.ref_path, rs.qry_path, sort_by_ref=True)


def showCoorMummer(specialForRaw, mummerLink, folderName, outputName, specialName):
    """alignerRobot.py:66-72: `show-coords -r <folder><out>.delta > <folder><out>Out | <folder><specialName>`."""
    _join_writes()
    out = folderName + (specialName if specialForRaw else outputName + "Out")
    _write_coords(_rowset_for(folderName + outputName + ".delta"), out)


def _align_job(folderName, outputName, referenceName, queryName, specialForRaw, specialName, refinedVersion, E=None, index_entry=None):
    """nucmer + show-coords of one job without waiting for its files: the .delta and the coords table are formatted and
    written on the writer pool while the caller goes on to the next job.  Callers join with _join_writes()."""
    qry = queryName if specialForRaw else folderName + queryName
    rs = _run_nucmer(folderName + outputName, folderName + referenceName, qry, _params(refinedVersion), E=E, index_entry=index_entry)
    out = folderName + (specialName if specialForRaw else outputName + "Out")
    if rs is None:
        open(out, "w").close()
    else:
        _submit((E or get_engine()).write_coords, out, rs.rows, rs.ref_names, rs.qry_names, rs.ref_path, rs.qry_path, True)
    return rs


_SUPER_BASES = 256_000_000     # query bases per GPU call when jobs are merged.  A 10 kb noisy read keeps one warp busy for
                               # milliseconds, so a 12.5 Mbp part alone fills a quarter of the resident warps, and every call ends
                               # on the tail of its slowest reads: measured on cfg 2 (250 Mbp in 20 parts, tools/super_batch_sweep.py)
                               # 32 Mbp per call 247 ms, 64: 189, 96: 186, 128: 204, one call: 178 ms -- the tails cost more than
                               # overlapping the upload of the next call saves.  Larger inputs still go in several calls.


def _merge_queries(entries):
    """entries: [(names, bases, off), ...] -> (bases, off) of their concatenation; adjacent views into one parent array are
    merged without a copy (the parts handed over by the splitter are)."""
    if len(entries) == 1:
        return entries[0][1], entries[0][2]
    bufs = [e[1] for e in entries]
    adjacent = all(isinstance(b, np.ndarray) and b.base is not None and b.base is bufs[0].base for b in bufs)
    if adjacent:
        addr = [b.__array_interface__["data"][0] for b in bufs]
        adjacent = all(addr[i] + bufs[i].size == addr[i + 1] for i in range(len(bufs) - 1))
    total = sum(int(b.size) for b in bufs)
    if adjacent:
        parent = bufs[0].base
        start = addr[0] - parent.__array_interface__["data"][0]
        merged = parent[start:start + total] if parent.dtype == np.uint8 and parent.ndim == 1 else np.concatenate(bufs)
    else:
        merged = np.concatenate(bufs)
    offs, acc = [np.zeros(1, dtype=np.int64)], 0
    for e in entries:
        o = np.asarray(e[2], dtype=np.int64)
        offs.append(o[1:] - o[0] + acc)
        acc += int(o[-1] - o[0])
    return merged, np.concatenate(offs)


def _split_rows(rows, bounds):
    """Rows of a merged batch (ascending read index) -> one Rows per [b, e) read range, read ids rebased."""
    cut = np.searchsorted(rows.qry_id, [b for b, _ in bounds] + [bounds[-1][1]]) if len(rows) else np.zeros(len(bounds) + 1, dtype=np.int64)
    out = []
    for k, (b, e) in enumerate(bounds):
        lo, hi = int(cut[k]), int(cut[k + 1])
        r = _engine_mod.Rows()
        for f in ("ref_id", "s1", "e1", "s2", "e2", "errors", "cols"):
            setattr(r, f, np.ascontiguousarray(getattr(rows, f)[lo:hi]))
        r.qry_id = np.ascontiguousarray(rows.qry_id[lo:hi] - b)
        r.stats, r.ms, r.mems, r.clusters = rows.stats, rows.ms, None, None
        r.delta_begin = r.delta_end = r.deltas = None
        if rows.deltas is not None and hi > lo:
            db, de = rows.delta_begin[lo:hi], rows.delta_end[lo:hi]
            ok = db >= 0
            d0 = int(db[ok].min()) if ok.any() else 0
            d1 = int(de[ok].max()) if ok.any() else 0
            r.deltas = np.ascontiguousarray(rows.deltas[d0:d1])
            r.delta_begin = np.where(ok, db - d0, -1).astype(np.int64)
            r.delta_end = np.where(ok, de - d0, -1).astype(np.int64)
        elif rows.deltas is not None:
            r.deltas = np.zeros(0, dtype=np.int32)
            r.delta_begin = np.zeros(0, dtype=np.int64)
            r.delta_end = np.zeros(0, dtype=np.int64)
        out.append(r)
    return out


def _batch_serial(folderName, workerList, specialForRaw, refinedVersion):
    """The jobs of one batch call on one GPU.  Consecutive jobs that share the reference are merged into GPU calls of about
    _SUPER_BASES query bases (every read is aligned on its own, so the rows of a merged call split exactly into the rows of
    its jobs); the next merged batch is uploaded and packed by a second host thread while the current one is aligned, and
    the .delta / coords text of the finished jobs is written on the writer pool."""
    import queue
    E = get_engine()
    params = _params(refinedVersion)
    flags = _engine_mod.WANT_DELTAS if fullDelta else 0
    jobs = []
    for outputName, referenceName, queryName, specialName in workerList:
        qry = queryName if specialForRaw else folderName + queryName
        jobs.append(dict(prefix=folderName + outputName, ref=folderName + referenceName, qry=qry,
                         out=folderName + (specialName if specialForRaw else outputName + "Out")))
    k = 0
    while k < len(jobs):
        g = k
        while g < len(jobs) and jobs[g]["ref"] == jobs[k]["ref"]:
            g += 1
        group, k = jobs[k:g], g
        ref_path = group[0]["ref"]
        live = []
        for j in group:
            with _CACHE_LOCK:
                parsed = _QUERY_CACHE.get(os.path.abspath(j["qry"]))
            if not (os.path.exists(ref_path) and (parsed is not None or os.path.exists(j["qry"]))):
                print("ERROR: Could not find input file(s) %s %s" % (ref_path, j["qry"]))   # nucmer's failure is ignored upstream too
                delta = j["prefix"] + ".delta"
                with _CACHE_LOCK:
                    _ROWS_CACHE.pop(delta, None)
                _PENDING_ROWS.pop(delta, None)
                if os.path.exists(delta):
                    os.remove(delta)
                open(j["out"], "w").close()
                continue
            j["parsed"] = parsed
            live.append(j)
        if not live:
            continue
        ix, rnames, rlens = _get_index(E, ref_path, params)
        dev = ix.info()["device"]
        # merged batches
        batches, cur, cur_bases = [], [], 0
        for j in live:
            if j["parsed"] is None:
                j["parsed"] = fasta.read_fasta_arrays(j["qry"])
            nb = int(j["parsed"][2][-1] - j["parsed"][2][0])
            if cur and cur_bases + nb > _SUPER_BASES:
                batches.append(cur)
                cur, cur_bases = [], 0
            cur.append(j)
            cur_bases += nb
        if cur:
            batches.append(cur)
        q = queue.Queue(maxsize=2)

        def uploader():
            E2 = None
            try:
                E2 = _engine_mod.Engine(E.L)
                E2.set_device(dev)
                for b in batches:
                    buf, off = _merge_queries([j["parsed"] for j in b])
                    q.put((b, E2.upload((buf, off)) if len(off) > 1 else None, None))
            except Exception as e:        # surfaced in the caller's thread
                q.put((None, None, e))
            finally:
                if E2 is not None:
                    E2.thread_release()
        th = threading.Thread(target=uploader)
        th.start()
        try:
            for _ in batches:
                b, reads, err = q.get()
                if err is not None:
                    raise err
                if reads is None:
                    rows = E.align(ix, [], params)
                else:
                    try:
                        rows = E.align_reads(ix, reads, params, flags)
                    except _engine_mod._lib.MfscError as e:
                        if not fullDelta or "delta lists" not in str(e):
                            raise
                        print("note: delta lists did not fit, writing alignment headers only (%s)" % e)
                        rows = E.align_reads(ix, reads, params, 0)
                    finally:
                        reads.free()
                bounds, acc = [], 0
                for j in b:
                    n = len(j["parsed"][2]) - 1
                    bounds.append((acc, acc + n))
                    acc += n
                for j, r in zip(b, _split_rows(rows, bounds)):
                    qnames = [n.split()[0] if n.split() else n for n in j["parsed"][0]]
                    qlens = np.diff(j["parsed"][2]).tolist()
                    rs = RowSet(r, rnames, rlens, qnames, qlens, ref_path, j["qry"])
                    _finish_job(E, j["prefix"], j["out"], rs)
        finally:
            th.join()
    _join_writes()


def _finish_job(E, prefix, out, rs):
    """The files of one finished job, on the writer pool: <prefix>.delta, then the coords table."""
    delta = prefix + ".delta"
    with _CACHE_LOCK:
        _ROWS_CACHE.pop(delta, None)
    _PENDING_ROWS[delta] = rs

    def write():
        E.write_delta(delta, rs.rows, rs.ref_names, rs.ref_lens, rs.qry_names, rs.qry_lens, os.path.abspath(rs.ref_path), os.path.abspath(rs.qry_path))
        _rows_cache_put(delta, rs)
        E.write_coords(out, rs.rows, rs.ref_names, rs.qry_names, rs.ref_path, rs.qry_path, True)
    _submit(write)


def useMummerAlign(mummerLink, folderName, outputName, referenceName, queryName, specialForRaw=False, specialName="",
                   refinedVersion=False):
    _align_job(folderName, outputName, referenceName, queryName, specialForRaw, specialName, refinedVersion)
    _join_writes()


def zeropadding(i):
    return "0" + str(i) if i < 10 else str(i)


def combineMultipleCoorMum(specialForRaw, mummerLink, folderName, outputName, specialName, numberOfFiles):
    """alignerRobot.py:74-96: show-coords every part, then `head -5` of part 01 + `tail -n+6` of every part."""
    print("")
    for i in range(1, numberOfFiles + 1):
        showCoorMummer(specialForRaw, mummerLink, folderName, outputName + zeropadding(i), specialName + zeropadding(i))
    # note: with specialForRaw False showCoorMummer writes <out>NNOut, and the reference then reads <specialName>NN
    with open(folderName + specialName, "w") as dst:
        with open(folderName + specialName + "01") as f:
            dst.writelines(f.readlines()[:5])
        for i in range(1, numberOfFiles + 1):
            with open(folderName + specialName + zeropadding(i)) as f:
                dst.writelines(f.readlines()[5:])


def useMummerAlignBatch(mummerLink, folderName, workerList, nProc, specialForRaw=False, refinedVersion=False):
    """alignerRobot.py:114-236.  workerList: [[outputName, referenceName, queryName, specialName], ...].
    The Pool / MPI fan-out becomes a loop over the jobs on the GPU (nProc is advisory); jobs that share a
    reference share its index.  globalLarge reproduces the row order of the 10x10 tiling."""
    if not houseKeeper.globalLarge:
        if houseKeeper.globalGPUs > 1 and len(workerList) > 1:
            _batch_on_gpus(folderName, workerList, specialForRaw, refinedVersion, houseKeeper.globalGPUs)
            return
        global _BATCH_SIGS
        _BATCH_SIGS = {}
        try:
            _batch_serial(folderName, workerList, specialForRaw, refinedVersion)
        finally:
            _BATCH_SIGS = None
        return
    # globalLarge (alignerRobot.py:168-236): reference and query are each split into 10 parts (written to the CWD, where
    # fasta-splitter.pl puts them), every (reference part i, query part j) tile is its own nucmer job
    # `<folder><out>IIJJ.delta`, and the coords file is `head -5` of tile (01, 01) followed by `tail -n+6` of every tile,
    # i outer, j inner.  The tiles are run for real: a cluster that spans two reference records (mgaps chains over the
    # concatenated reference) exists only when both records are in the same part, so re-ordering the rows of one whole
    # alignment would not give the same table.
    n_parts = 10
    params = _params(refinedVersion)
    for outputName, referenceName, queryName, specialName in workerList:
        qry_full = queryName if specialForRaw else folderName + queryName
        for path in (folderName + referenceName, qry_full):
            if os.path.exists(path):
                fasta.split_fasta(path, n_parts, out_dir=os.getcwd())
        out = folderName + (specialName if specialForRaw else outputName + "Out")
        lines, header = [], None
        for i in range(1, n_parts + 1):
            for j in range(1, n_parts + 1):
                tmpRef = referenceName[0:-6] + ".part-" + zeropadding(i) + ".fasta"
                tmpQry = queryName[0:-6] + ("-" if specialForRaw else ".part-") + zeropadding(j) + ".fasta"
                prefix = folderName + outputName + zeropadding(i) + zeropadding(j)
                rs = _run_nucmer(prefix, tmpRef, tmpQry, params)
                tile_out = prefix + ".coords.tmp"
                _write_coords(rs, tile_out)
                with open(tile_out) as f:
                    tl = f.readlines()
                os.remove(tile_out)
                if i == 1 and j == 1:
                    header = tl[:5]
                lines += tl[5:]
        with open(out, "w") as f:
            f.writelines((header or []) + lines)
    _join_writes()


def _batch_on_gpus(folderName, workerList, specialForRaw, refinedVersion, n_gpus):
    """The reference's Pool(nProc) fan-out (alignerRobot.py:154-166) over the GPUs of one box: every distinct reference
    of the batch is indexed ONCE (on the calling thread's device, through the index cache) and replicated to the other
    GPUs with one NCCL broadcast per buffer (mfsc_index_broadcast); one host thread per GPU then takes the jobs
    round-robin (the C library keeps one device context and stream per thread).  Output files are those of the serial loop."""
    from . import multigpu
    main = get_engine()
    params = _params(refinedVersion)
    errors = []
    replicas = {}                                # reference path -> ([Index per device], names, lens)
    clique = None
    try:
        for job in workerList:
            ref_path = folderName + job[1]
            if ref_path in replicas or not os.path.exists(ref_path):
                continue
            ix, names, lens = _get_index(main, ref_path, params)
            if clique is None:
                root_dev = ix.info()["device"]
                devices = [root_dev] + [d for d in range(n_gpus) if d != root_dev]
                clique = multigpu.LocalClique(main, devices)
            per_dev, _ = clique.broadcast(ix, root=0)
            replicas[ref_path] = (per_dev, names, lens)
        devices = clique.devices if clique is not None else list(range(n_gpus))

        def work(slot):
            E = None
            try:
                E = _engine_mod.Engine(main.L)
                E.set_device(devices[slot])
                for k in range(slot, len(workerList), n_gpus):
                    outputName, referenceName, queryName, specialName = workerList[k]
                    ref_path = folderName + referenceName
                    qry_path = queryName if specialForRaw else folderName + queryName
                    rep = replicas.get(ref_path)
                    entry = (rep[0][slot], rep[1], rep[2]) if rep is not None else None
                    _align_job(folderName, outputName, referenceName, queryName, specialForRaw, specialName, refinedVersion, E=E, index_entry=entry)
            except Exception as e:       # surfaced in the caller's thread
                errors.append(e)
            finally:
                if E is not None:
                    E.thread_release()   # the worker's stream goes with the worker

        threads = [threading.Thread(target=work, args=(d,)) for d in range(n_gpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        _join_writes()
    finally:
        for per_dev, _, _ in replicas.values():
            for ix in per_dev[1:]:               # entry 0 is the cached index of the calling thread
                ix.free()
        if clique is not None:
            clique.free()
    if errors:
        raise errors[0]


def transformCoor(dataList):
    """alignerRobot.py:242-256: order S2/E2 ascending (in place, like the reference)."""
    newList = []
    for eachitem in dataList:
        if not eachitem[2] < eachitem[3]:
            eachitem[2], eachitem[3] = eachitem[3], eachitem[2]
        newList.append(eachitem)
    return newList


def extractMumData(folderName, fileName):
    """alignerRobot.py:7-40: skip 5 lines, then per row [S1, E1, S2, E2, LEN1, LEN2, IDY, refName, qryName]."""
    dataList = []
    with open(folderName + fileName, "r") as f:
        lines = f.readlines()
    for tmp in lines[5:]:
        tmp = tmp.rstrip()
        if len(tmp) == 0:
            break
        info = tmp.split("|")
        first, second, third = info[0].split(), info[1].split(), info[2].split()
        names = info[-1].split("\t")
        dataList.append([int(first[0]), int(first[1]), int(second[0]), int(second[1]), int(third[0]), int(third[1]),
                         float(info[3]), names[0].strip(), names[1].strip()])
    return dataList


def largeRvsQAlign(folderName, numberOfFiles, mummerLink, refFile, qryFile, mumTmpHeader):
    """alignerRobot.py:258-297: split the query FASTA into parts, align every part against the reference,
    return the concatenated rows of parts 01..NN (each part in show-coords -r order).
    Pipeline: the query file is parsed once (a few host threads); the part files the rest of the reference pipeline
    re-reads are written on the writer pool while the parts, handed over in memory, are aligned one after the other on
    the GPU; the .delta and coords text of part i is formatted and written while part i+1 is aligned."""
    parts = fasta.split_fasta(folderName + qryFile + ".fasta", numberOfFiles, out_dir=folderName, submit=_submit, keep=_QUERY_CACHE)
    try:
        numberOfFiles = houseKeeper.globalParallelFileNum      # the reference re-reads the global here (alignerRobot.py:271)
        workerList = []
        for i in range(1, numberOfFiles + 1):
            idx = zeropadding(i)
            workerList.append([mumTmpHeader + idx, refFile + ".fasta", qryFile + ".part-" + idx + ".fasta", mumTmpHeader + idx])
        useMummerAlignBatch(mummerLink, folderName, workerList, houseKeeper.globalParallel, False)
    finally:
        with _CACHE_LOCK:
            for p in parts:
                _QUERY_CACHE.pop(os.path.abspath(p), None)
    dataList = []
    for i in range(1, numberOfFiles + 1):
        rs = _rows_cache_get(folderName + mumTmpHeader + zeropadding(i) + ".delta")
        # rows served from memory when this process produced them; identical to parsing the text
        dataList += rs.datalist() if rs is not None else extractMumData(folderName, mumTmpHeader + zeropadding(i) + "Out")
    return dataList
