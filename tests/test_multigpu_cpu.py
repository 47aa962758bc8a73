"""World-size-2 test of the N>1 host logic on CPU (gloo): read batches shard by contiguous range across
ranks with no data-path collective, each rank aligns its shard (emulation build here), rows gathered in
rank order equal the single-process rows, and the max-over-ranks timing reduction works."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, pickle
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from metafinishersc_b200 import _lib, engine, multigpu, datagen
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    E = engine.Engine(_lib.bind(os.path.join(%r, "tests", "emu", "_build", "libmfsc_emu.so")))
    d = datagen.abun_gap(G=60_000, repeat_len=1200, L=3000, seed=4)
    reads = d["reads"][:90]
    P = engine.make_params()
    ix = E.index(d["contigs"], P)                       # on GPUs: multigpu.replicated_index (NCCL broadcast)
    b, e = multigpu.shard_reads(len(reads), rank, world)
    r = E.align(ix, reads[b:e], P)
    mine = np.stack([r.ref_id, r.qry_id + b, r.s1, r.e1, r.s2, r.e2, r.errors, r.cols], axis=1)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        full = E.align(ix, reads, P)
        want = np.stack([full.ref_id, full.qry_id, full.s1, full.e1, full.s2, full.e2, full.errors, full.cols], axis=1)
        got = np.concatenate(gathered)
        assert np.array_equal(got, want), "sharded rows differ from the single-process rows"
        print("MULTI_OK", len(got))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path, emu_engine):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_OK" in out.stdout


def test_shard_ranges():
    from metafinishersc_b200 import multigpu
    for n in (0, 1, 7, 20, 25000):
        for w in (1, 2, 4, 8):
            r = [multigpu.shard_reads(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
