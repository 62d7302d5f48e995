"""Frame sharding across processes and the hand-over of the communicator id.

Frames are independent units for every accumulator on the path (density, rdf, orient are sums over frames of
per-frame integer counts: reference README.md:148, calc_rdf_region.cpp:91, calc_orient_mof.cpp:180,196), so each
process owns a contiguous block of frames and private u64 histograms, and ONE all-reduce (sum) merges them at the
end.  The collective itself lives in libb2a (`b2a_finalize`: ncclAllReduce of every accumulator, ncclAllGather of
the per-frame series; include/b2a.h): this module only does what a launcher has to do -- split the frame range and
carry the 128-byte NCCL id from rank 0 to the others over whatever channel the launcher provides
(`torch.distributed` here: NCCL under torchrun on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_frames(nframes: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`: GPU g gets frames [g*F/G, (g+1)*F/G)."""
    lo = (nframes * rank) // world
    hi = (nframes * (rank + 1)) // world
    return lo, hi


def broadcast_id(make_id, device=None) -> bytes:
    """Rank 0 calls `make_id()` (128 bytes); everybody returns rank 0's bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def comm_init(ctx, device=None):
    """Joins `ctx` (one process = one context, one or several GPUs) to a communicator spanning every process of the
    torch.distributed world.  No-op outside a multi-process launch."""
    import torch.distributed as dist
    from . import lib
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    uid = broadcast_id(lib.comm_unique_id, device)
    ctx.comm_init(dist.get_world_size(), dist.get_rank(), uid)
    return True


def allreduce_host_counts(counts: np.ndarray) -> np.ndarray:
    """Host-side sum over ranks (gloo CPU tests of the sharding arithmetic; the product merges on the devices)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts).view(np.int64).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint64)
