"""Several GPUs of one box: the contig index is built ONCE and replicated with one NCCL broadcast per buffer over
NVLink (`mfsc_index_broadcast` in the C ABI); reads shard by batch, so alignment itself needs no collective
(the reference's analogue is the 20-part fan-out of alignerRobot.useMummerAlignBatch,
alignerRobot.py:114-166, where every job rebuilds the suffix tree of the whole reference).

Two launch shapes:
  * one process, one host thread per GPU (`LocalClique`): what alignerRobot._batch_on_gpus uses;
  * one process per GPU under torchrun (`RankComm`): the 128-byte NCCL id travels through torch.distributed
    (plumbing), the broadcast itself is the library's.
"""
import ctypes

import numpy as np

from . import _lib
from .engine import Index


class LocalClique:
    """NCCL communicators over `devices` of this process (ncclCommInitAll)."""

    def __init__(self, E, devices):
        self.L, self.devices = E.L, list(devices)
        dv = np.ascontiguousarray(self.devices, dtype=np.int32)
        h = ctypes.c_void_p()
        _lib.check(self.L, self.L.mfsc_comm_init_local(_lib.ptr(dv), len(dv), ctypes.byref(h)))
        self.h = h

    def broadcast(self, index, root=0):
        """index: the built Index living on devices[root].  Returns ([Index per device], ms); entry `root` is `index`."""
        n = len(self.devices)
        arr = (ctypes.c_void_p * n)()
        arr[root] = index.h
        ms = ctypes.c_double()
        _lib.check(self.L, self.L.mfsc_index_broadcast(self.h, root, arr, ctypes.byref(ms)))
        out = [index if i == root else Index(self.L, ctypes.c_void_p(arr[i]), index.n_rec) for i in range(n)]
        return out, ms.value

    def free(self):
        if self.h:
            self.L.mfsc_comm_free(self.h)
            self.h = None


class RankComm:
    """One rank of a one-process-per-GPU job.  `exchange(buf, src)` must broadcast a uint8 numpy array of 128 bytes from
    rank `src` to all ranks in place (torch.distributed in bench.py; anything else works as well)."""

    def __init__(self, E, rank, world, exchange):
        self.L, self.rank, self.world = E.L, rank, world
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(self.L, self.L.mfsc_comm_unique_id(_lib.ptr(uid)))
        exchange(uid, 0)
        h = ctypes.c_void_p()
        _lib.check(self.L, self.L.mfsc_comm_init_rank(_lib.ptr(uid), rank, world, ctypes.byref(h)))
        self.h = h

    def broadcast(self, index, n_rec, root=0):
        """index: the built Index on the root rank, None elsewhere.  Returns (Index, ms)."""
        arr = (ctypes.c_void_p * 1)()
        arr[0] = index.h if index is not None else None
        ms = ctypes.c_double()
        _lib.check(self.L, self.L.mfsc_index_broadcast(self.h, root, arr, ctypes.byref(ms)))
        if index is None:
            index = Index(self.L, ctypes.c_void_p(arr[0]), n_rec)
        return index, ms.value

    def build_sharded(self, contigs, params):
        """Every rank calls it with the same contigs: rank r builds 1/N of the k-mer space, the slices are exchanged over
        NVLink (mfsc_index_build_sharded).  Returns (Index, [upload, own slice, exchange, total] ms)."""
        from .engine import _as_buf
        buf, off = _as_buf(contigs)
        h = ctypes.c_void_p()
        ms = np.zeros(4, dtype=np.float64)
        _lib.check(self.L, self.L.mfsc_index_build_sharded(self.h, _lib.ptr(buf), _lib.ptr(off), len(off) - 1, ctypes.byref(params),
                                                           ctypes.byref(h), _lib.ptr(ms)))
        return Index(self.L, h, len(off) - 1), ms.tolist()

    def free(self):
        if self.h:
            self.L.mfsc_comm_free(self.h)
            self.h = None


def replicated_index(E, contigs, params, rank, world, comm=None):
    """Rank 0 builds, every rank receives (RankComm).  Returns (index, broadcast_ms).  world == 1: plain build."""
    if world == 1:
        return E.index(contigs, params), 0.0
    ix = E.index(contigs, params) if rank == 0 else None
    n_rec = len(contigs[1]) - 1 if isinstance(contigs, tuple) else len(contigs)
    return comm.broadcast(ix, n_rec, root=0)


def shard_reads(n_reads, rank, world):
    """Contiguous read range of a rank (batches keep the reference's part order)."""
    per = (n_reads + world - 1) // world
    return min(n_reads, rank * per), min(n_reads, (rank + 1) * per)


def gather_rows(rows_list):
    """Concatenate per-rank row sets (already in rank = part order)."""
    out = {}
    for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
        out[f] = np.concatenate([getattr(r, f) for r in rows_list]) if rows_list else np.zeros(0, dtype=np.int32)
    return out
