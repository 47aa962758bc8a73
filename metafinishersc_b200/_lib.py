"""ctypes binding of libmfsc_b200.so (the C ABI in include/mfsc.h).

The product library is `csrc/libmfsc_b200.so`, built in-tree by `build.py` / `__graft_entry__.build()`.
`load()` raises if it is missing -- there is no CPU fallback and nothing in this package computes
alignments any other way.  (`bind(path)` exists so that tests can bind the host-emulation build of
the same sources under tests/emu/; the package itself never calls it with anything but the product
library.)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmfsc_b200.so")
_DEV = os.environ.get("MFSC_LIB_DEV")        # developer hook: A/B another nvcc build of the SAME CUDA library (tools/, build_ab/)
if _DEV:
    _real = os.path.realpath(_DEV)
    _root = os.path.dirname(_HERE)
    for _bad in (os.path.join(_root, "oracle"), os.path.join(_root, "tests")):
        if _real.startswith(os.path.realpath(_bad) + os.sep):
            raise ImportError("MFSC_LIB_DEV must name a CUDA build of libmfsc_b200, not test infrastructure (%s)" % _real)
    if "emu" in os.path.basename(_real) or "oracle" in os.path.basename(_real):
        raise ImportError("MFSC_LIB_DEV must name a CUDA build of libmfsc_b200, not test infrastructure (%s)" % _real)
    LIB_PATH = _real

c_i32, c_i64, c_vp, c_dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_double


class Params(ctypes.Structure):
    """mfsc_params; defaults are nucmer's (alignerRobot.py:46-64)."""
    _fields_ = [("minmatch", c_i32), ("mincluster", c_i32), ("maxgap", c_i32), ("breaklen", c_i32), ("diagdiff", c_i32),
                ("maxmatch", c_i32), ("simplify", c_i32), ("band_cap", c_i32), ("diagfactor", c_dbl), ("kmer", c_i32),
                ("engine", c_i32), ("pk_u_span", c_i32), ("pk_m_span", c_i32)]


def make_params(minmatch=20, mincluster=65, maxgap=90, breaklen=200, diagdiff=5, diagfactor=0.12, maxmatch=True,
                simplify=True, band_cap=248, kmer=0, engine=0, pk_u_span=0, pk_m_span=0):
    p = Params()
    p.minmatch, p.mincluster, p.maxgap, p.breaklen, p.diagdiff = minmatch, mincluster, maxgap, breaklen, diagdiff
    p.maxmatch, p.simplify, p.band_cap, p.diagfactor, p.kmer = int(maxmatch), int(simplify), band_cap, diagfactor, kmer
    p.engine, p.pk_u_span, p.pk_m_span = engine, pk_u_span, pk_m_span
    return p


KEEP_MEMS, KEEP_CLUSTERS, STOP_SEEDS, STOP_CLUSTERS, WANT_DELTAS = 1, 2, 4, 8, 16
# mfsc_row_flags bits (include/mfsc.h)
(RF_HEADTAIL, RF_IDENTICAL, RF_ASSOCIATED, RF_EMBEDDED_LR, RF_REF_COVERED, RF_REF_HEAD, RF_SHORT_REF, RF_SHORT_QRY, RF_DIFF_NAME,
 RF_END_SPAN, RF_OVERLAP) = [1 << i for i in range(11)]

# every symbol include/mfsc.h declares (tests check the library exports all of them)
SYMBOLS = ["mfsc_last_error", "mfsc_version", "mfsc_device_count", "mfsc_set_device", "mfsc_thread_release", "mfsc_trim_pool", "mfsc_default_params",
           "mfsc_host_alloc", "mfsc_host_free", "mfsc_index_build", "mfsc_index_create", "mfsc_index_info",
           "mfsc_index_buffers", "mfsc_index_build_ms", "mfsc_index_free", "mfsc_comm_unique_id", "mfsc_comm_init_rank",
           "mfsc_comm_init_local", "mfsc_index_broadcast", "mfsc_comm_free", "mfsc_index_build_sharded", "mfsc_reads_upload", "mfsc_reads_free",
           "mfsc_align_reads", "mfsc_align_batch", "mfsc_result_count", "mfsc_result_rows", "mfsc_result_mems",
           "mfsc_result_clusters", "mfsc_result_deltas", "mfsc_result_stats", "mfsc_result_free", "mfsc_coverage", "mfsc_cov_build", "mfsc_cov_fetch", "mfsc_cov_interval_sums", "mfsc_cov_free", "mfsc_row_flags", "mfsc_write_coords",
           "mfsc_write_delta", "mfsc_fasta_parse", "mfsc_fasta_count", "mfsc_fasta_write"]


class MfscError(RuntimeError):
    pass


def bind(path):
    L = ctypes.CDLL(path)
    L.mfsc_last_error.restype = ctypes.c_char_p
    L.mfsc_version.restype = c_i32
    L.mfsc_result_count.restype = c_i64
    L.mfsc_result_count.argtypes = [c_vp, c_i32]
    PP = ctypes.POINTER(Params)
    L.mfsc_device_count.argtypes = [ctypes.POINTER(c_i32)]
    L.mfsc_set_device.argtypes = [c_i32]
    L.mfsc_default_params.argtypes = [PP]
    L.mfsc_host_alloc.argtypes = [ctypes.POINTER(c_vp), c_i64]
    L.mfsc_host_free.argtypes = [c_vp]
    L.mfsc_index_build.argtypes = [c_vp, c_vp, c_i32, PP, ctypes.POINTER(c_vp)]
    L.mfsc_index_create.argtypes = [c_vp, c_vp, c_i32, PP, c_i32, c_i64, ctypes.POINTER(c_vp)]
    L.mfsc_index_info.argtypes = [c_vp, c_vp]
    L.mfsc_index_buffers.argtypes = [c_vp, c_vp, c_vp]
    L.mfsc_index_build_ms.argtypes = [c_vp, c_vp]
    L.mfsc_index_free.argtypes = [c_vp]
    L.mfsc_comm_unique_id.argtypes = [c_vp]
    L.mfsc_comm_init_rank.argtypes = [c_vp, c_i32, c_i32, ctypes.POINTER(c_vp)]
    L.mfsc_comm_init_local.argtypes = [c_vp, c_i32, ctypes.POINTER(c_vp)]
    L.mfsc_index_broadcast.argtypes = [c_vp, c_i32, c_vp, ctypes.POINTER(c_dbl)]
    L.mfsc_comm_free.argtypes = [c_vp]
    L.mfsc_index_build_sharded.argtypes = [c_vp, c_vp, c_vp, c_i32, PP, ctypes.POINTER(c_vp), c_vp]
    L.mfsc_reads_upload.argtypes = [c_vp, c_vp, c_i32, ctypes.POINTER(c_vp)]
    L.mfsc_reads_free.argtypes = [c_vp]
    L.mfsc_align_reads.argtypes = [c_vp, c_vp, PP, c_i32, ctypes.POINTER(c_vp)]
    L.mfsc_align_batch.argtypes = [c_vp, c_vp, c_vp, c_i32, PP, c_i32, ctypes.POINTER(c_vp)]
    L.mfsc_result_rows.argtypes = [c_vp] * 9
    L.mfsc_result_mems.argtypes = [c_vp] * 6
    L.mfsc_result_clusters.argtypes = [c_vp] * 9
    L.mfsc_result_deltas.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.mfsc_result_stats.argtypes = [c_vp, c_vp, c_vp]
    L.mfsc_result_free.argtypes = [c_vp]
    L.mfsc_coverage.argtypes = [c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp, c_vp]
    L.mfsc_cov_build.argtypes = [c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, ctypes.POINTER(c_vp), c_vp]
    L.mfsc_cov_fetch.argtypes = [c_vp, c_i32, c_vp]
    L.mfsc_cov_interval_sums.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]
    L.mfsc_cov_free.argtypes = [c_vp]
    L.mfsc_row_flags.argtypes = [c_i64] + [c_vp] * 7 + [c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]
    L.mfsc_write_coords.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, c_i64] + [c_vp] * 8 + \
        [ctypes.c_char_p, c_i32, ctypes.c_char_p, c_i32, c_i32]
    L.mfsc_write_delta.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, c_i64] + [c_vp] * 8 + \
        [ctypes.c_char_p, c_vp, c_i32, ctypes.c_char_p, c_vp, c_i32, c_vp, c_vp, c_vp]
    L.mfsc_fasta_parse.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, ctypes.POINTER(c_i32)]
    L.mfsc_fasta_count.argtypes = [c_vp, c_i64, ctypes.POINTER(c_i32)]
    L.mfsc_fasta_write.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32]
    return L


_lib = None


def load():
    """The product library.  Raises when it has not been built: no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MfscError("libmfsc_b200.so is not built (run `python -m metafinishersc_b200.build`); "
                            "this package has no CPU fallback")
        _lib = bind(LIB_PATH)
    return _lib


def check(L, rc):
    if rc != 0:
        raise MfscError("libmfsc rc=%d: %s" % (rc, (L.mfsc_last_error() or b"").decode()))


def ptr(a):
    return a.ctypes.data if a is not None else None


def concat(seqs):
    """list[bytes] -> (uint8 array, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=off[1:])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, dtype=np.uint8)
    return buf, off
