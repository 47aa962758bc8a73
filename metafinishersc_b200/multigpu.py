"""One process per GPU: the contig index is built on rank 0 and replicated with one NCCL broadcast
per buffer over NVLink (torch.distributed); reads shard by batch, so alignment itself needs no
collective (the reference's analogue is the 20-part fan-out of
alignerRobot.useMummerAlignBatch, alignerRobot.py:114-166).
"""
import time

import numpy as np


class _DevBuf:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def replicated_index(E, contigs, params, rank, world):
    """Returns (index, broadcast_ms).  world == 1: plain build."""
    if world == 1:
        return E.index(contigs, params), 0.0
    import torch
    import torch.distributed as dist
    meta = torch.zeros(2, dtype=torch.int64, device="cuda")
    ix = None
    if rank == 0:
        ix = E.index(contigs, params)
        info = ix.info()
        meta[0], meta[1] = info["k"], info["positions"]
    dist.broadcast(meta, src=0)
    if rank != 0:
        params.kmer = int(meta[0].item())
        ix = E.index(contigs, params, build_tables=False, n_positions=int(meta[1].item()))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for ptr, nbytes in ix.buffers():
        t = torch.as_tensor(_DevBuf(ptr, nbytes), device="cuda")
        dist.broadcast(t, src=0)
    torch.cuda.synchronize()
    return ix, (time.perf_counter() - t0) * 1e3


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
