"""Batch sharding for the multi-GPU path.

Every sample -- and its matrix gradient -- is independent, so the batch dimension shards with no
data-path collective: each rank works on ``shard_range(B, world, rank)``.  When the data starts on
one rank, :func:`sharded_apply` wraps the op in ONE scatter of the inputs and ONE gather of the
outputs (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors for the tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced chunk [start, stop) of ``batch`` samples owned by ``rank``; the first
    ``batch % world`` ranks get one extra sample."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(batch: int, world: int) -> list[int]:
    return [shard_range(batch, world, r)[1] - shard_range(batch, world, r)[0] for r in range(world)]


def _scatter_rows(full: Optional[torch.Tensor], like_shape: Sequence[int], dtype, device, batch: int, src: int, group) -> torch.Tensor:
    """Scatter dim-0 chunks of ``full`` (present on ``src`` only).  Chunks are padded to a common
    length because the collective needs equal sizes; the padding is trimmed on arrival."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(batch, world)
    longest = max(sizes)
    recv = torch.empty((longest,) + tuple(like_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = shard_range(batch, world, r)
            c = full[a:b]
            if c.shape[0] < longest:
                c = torch.cat([c, c.new_zeros((longest - c.shape[0],) + tuple(like_shape))])
            chunks.append(c.contiguous())
    dist.scatter(recv, chunks, src=src, group=group)
    return recv[: sizes[rank]]


def _gather_rows(local: torch.Tensor, batch: int, dst: int, group) -> Optional[torch.Tensor]:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(batch, world)
    longest = max(sizes)
    padded = local
    if local.shape[0] < longest:
        padded = torch.cat([local, local.new_zeros((longest - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)])


def sharded_apply(fn: Callable[..., torch.Tensor], batched: Sequence[Optional[torch.Tensor]], *, batch: int, shapes, dtypes, device,
                  root: int = 0, group=None) -> Optional[torch.Tensor]:
    """``fn(*local_shards)`` on every rank, with the batched inputs scattered from ``root`` and the
    result gathered back to it.  ``batched`` holds the full tensors on ``root`` (anything on other
    ranks); ``shapes``/``dtypes`` describe one sample of each so non-root ranks can allocate."""
    locals_ = [_scatter_rows(t, s, d, device, batch, root, group) for t, s, d in zip(batched, shapes, dtypes)]
    out = fn(*locals_)
    return _gather_rows(out, batch, root, group)
