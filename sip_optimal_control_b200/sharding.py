"""Batch sharding across ranks (one process per GPU).

Problems are independent, so the batch is cut into contiguous
slices with no data-path collective.  The only exchange per Newton iteration
is the all-reduce of the 4-slot statistics vector the residual / status kernels
produce: {sum of squared residual norms, max residual norm, #failed, #problems}.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous slice; slices differ by at most one problem."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_stats(stats, group=None):
    """In-place all-reduce of a [4] float64 tensor: slots 0, 2, 3 are summed, slot 1
    (max residual norm) is max-reduced.  Works on any backend (nccl / gloo)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    mx = stats[1:2].clone()
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    stats[1:2] = mx
    return stats
