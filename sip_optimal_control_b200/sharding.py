"""Batch sharding across ranks (one process per GPU).

Problems are independent, so the batch is cut into contiguous
slices with no data-path collective.  The only exchange per Newton iteration
is the all-reduce of the 4-slot statistics vector the residual / status kernels
produce: {sum of squared residual norms, max residual norm, #failed, #problems}.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple


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


class Communicator:
    """One rank's endpoint of the engine's own NCCL clique (include/sipoc.h, "several
    devices"): ``allreduce_stats`` is the one exchange of a Newton iteration, and an engine it
    is attached to finishes every statistics vector it writes with that all-reduce on the
    call's stream.  ``torch.distributed`` (any backend) only carries the 128-byte unique id."""

    def __init__(self, device: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist

        from ._capi import SIPOC_COMM_ID_BYTES, lib

        self._lib = lib
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = torch.cuda.current_device() if device is None else int(device)
        ident = torch.zeros(SIPOC_COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * SIPOC_COMM_ID_BYTES)()
            rc = lib.sipoc_comm_unique_id(buf)
            if rc != 0:
                raise RuntimeError(f"sipoc_comm_unique_id failed ({rc}): is libnccl.so.2 loadable?")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if world > 1:
            backend = dist.get_backend(group)
            carrier = ident.cuda(dev) if backend == "nccl" else ident
            dist.broadcast(carrier, src=0, group=group)
            ident = carrier.cpu()
        raw = (ctypes.c_uint8 * SIPOC_COMM_ID_BYTES)(*ident.tolist())
        handle = ctypes.c_void_p()
        rc = lib.sipoc_comm_create(raw, rank, world, dev, ctypes.byref(handle))
        if rc != 0:
            raise RuntimeError(f"sipoc_comm_create failed ({rc})")
        self._handle, self.rank, self.world, self.device = handle, rank, world, dev

    def attach(self, engine) -> None:
        """Every `stats` output of ``engine`` is all-reduced from now on."""
        rc = self._lib.sipoc_attach_comm(engine._handle, self._handle)
        if rc != 0:
            raise RuntimeError(f"sipoc_attach_comm failed ({rc})")
        self._attached = getattr(self, "_attached", []) + [engine]

    def allreduce_stats(self, stats, stream=None) -> None:
        import torch

        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        rc = self._lib.sipoc_comm_allreduce_stats(self._handle, stats.data_ptr(), int(s.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"sipoc_comm_allreduce_stats failed ({rc})")

    def close(self) -> None:
        if getattr(self, "_handle", None):
            for e in getattr(self, "_attached", []):
                if getattr(e, "_handle", None):
                    self._lib.sipoc_attach_comm(e._handle, None)
            self._lib.sipoc_comm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
