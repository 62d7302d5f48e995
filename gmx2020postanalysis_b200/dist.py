"""Frame sharding across ranks and the single histogram merge.

Frames are independent units for every accumulator on the path (density, rdf, orient are sums over frames of
per-frame integer counts: reference README.md:148, calc_rdf_region.cpp:91, calc_orient_mof.cpp:180,196), so each
rank owns a contiguous block of frames and private u64 histograms, and ONE all-reduce (sum) merges them at the
end.  Integer sums make the merged result bit-identical for any number of ranks.  One process per GPU;
torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np


def shard_frames(nframes: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`: GPU g gets frames [g*F/G, (g+1)*F/G)."""
    lo = (nframes * rank) // world
    hi = (nframes * (rank + 1)) // world
    return lo, hi


class _CudaArrayView:
    """Minimal __cuda_array_interface__ holder so torch can alias a device buffer owned by libb2a."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def device_view(ctx, acc):
    """torch int64 tensor aliasing the accumulator's device words (counters + stats block).  Counts stay far
    below 2^63, so summing them as int64 is the same as summing uint64 (NCCL via torch has no uint64 sum)."""
    import torch
    from . import lib
    n = ctx.acc_size(acc) + lib.STAT_NSLOTS
    ptr = ctx.acc_device_ptr(acc)
    return torch.as_tensor(_CudaArrayView(ptr, n), device=f"cuda:{ctx.device}")


def allreduce_accumulator(ctx, acc):
    """The single collective of the path: sum the rank-private histograms in place over NVLink (NCCL)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    ctx.sync()                      # libb2a works on its own stream
    t = device_view(ctx, acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(ctx.device)


def allreduce_host_counts(counts: np.ndarray) -> np.ndarray:
    """Host-side variant (gloo CPU tests and the C++ multi-context path's reference semantics)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts).view(np.int64).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint64)
