"""Multi-process host logic of the batch-sharded path, on CPU with gloo (world size 2 and 3).
The per-rank op is the CPU oracle here (the CUDA op needs a GPU); what is under test is the shard
arithmetic and the scatter -> op -> gather plumbing, which is device agnostic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kornia_b200.sharding import shard_range, shard_sizes, sharded_apply


def test_shard_ranges_partition_the_batch():
    for B in (1, 2, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(B, world)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import kornia_restated as R

        torch.manual_seed(0)
        src = torch.rand(B, 3, 20, 24)
        M = torch.eye(3)[None].repeat(B, 1, 1) + 0.02 * torch.randn(B, 3, 3)
        M[:, 2, :2] *= 0.01
        full = (src, M) if rank == 0 else (None, None)
        out = sharded_apply(lambda s, m: R.warp_perspective(s, m, (18, 22)), full, batch=B, shapes=[(3, 20, 24), (3, 3)],
                            dtypes=[torch.float32, torch.float32], device="cpu", root=0)
        if rank == 0:
            whole = R.warp_perspective(src, M, (18, 22))
            q.put(bool(torch.equal(out, whole)))
        else:
            assert out is None
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        # bench.py's e2e reduction: a rank without a value (its pinned allocation failed) still joins the collective
        import bench

        assert bench.max_over_ranks_or_none(dist, 10.0 + rank, "ok", "cpu") == (10.0 + world - 1, "ok")
        got = bench.max_over_ranks_or_none(dist, None if rank == 1 else 5.0, "failed here" if rank == 1 else "ok", "cpu")
        assert got == (None, "failed here" if rank == 1 else "unavailable on another rank")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 5), (2, 8), (3, 7)])
def test_scatter_op_gather_equals_whole_batch(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
