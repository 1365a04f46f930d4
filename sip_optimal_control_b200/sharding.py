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


def allgather_stats(stats, gathered, group=None):
    """The per-iteration exchange as ONE collective: every rank's [4] float64 vector lands
    in `gathered` ([world, 4], same device).  Nothing is folded on the device; the reader
    (the host's convergence test, or fold_stats) combines the rows.  32 bytes per rank:
    latency-bound, so one collective instead of a SUM and a MAX all-reduce is what counts."""
    import torch.distributed as dist

    dist.all_gather_into_tensor(gathered.view(-1), stats, group=group)
    return gathered


def fold_stats(gathered):
    """[world, 4] -> [4]: slots 0, 2, 3 summed over ranks, slot 1 (max residual norm) max-ed."""
    out = gathered.sum(dim=0)
    out[1] = gathered[:, 1].max()
    return out


def allreduce_stats(stats, group=None):
    """In-place all-reduce of a [4] float64 tensor: slots 0, 2, 3 are summed, slot 1
    (max residual norm) is max-reduced.  Works on any backend (nccl / gloo)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    gathered = torch.empty((world, 4), dtype=stats.dtype, device=stats.device)
    allgather_stats(stats, gathered, group=group)
    stats.copy_(fold_stats(gathered))
    return stats
