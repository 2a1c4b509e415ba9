"""Data-parallel plumbing for bench.py: the batch of independent games shards across ranks with
no hot-path collective.  torch.distributed is used only to agree on the
timing (MAX over ranks), to sum the work done, and to gather results at the end.  Backend-agnostic
so the same code runs under NCCL on GPUs and under gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_bounds(global_batch: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_metrics(elapsed_ms: float, done: int, device=None):
    """(max elapsed over ranks, total work over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), int(done)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    c = torch.tensor([done], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), int(c.item())


def gather_trajectories(local, device=None):
    """All-gather of one per-rank result array (equal shapes) -> [world, ...] numpy array."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(local), device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None].cpu().numpy()
    out = torch.empty((dist.get_world_size(),) + tuple(t.shape), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(out.view(-1), t.reshape(-1))
    return out.cpu().numpy()
