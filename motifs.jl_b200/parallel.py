"""Multi-GPU plumbing (one process per GPU, torch.distributed): the scan shards sequences over ranks with no
data-path collective and sums the per-motif counts once at the end; training averages the filter
gradients with one all_reduce per step.  Backend is NCCL on GPUs; the same code runs under gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """contiguous block of items owned by `rank` (blocks differ by at most one item)."""
    return n_items * rank // world, n_items * (rank + 1) // world


def shard_groups(n_items: int, group: int, rank: int, world: int):
    """contiguous block of whole groups (e.g. batches of 6 sequences, model.jl:9) owned by `rank`; the trailing
    n_items % group items are dropped exactly like Flux.DataLoader(partial=false) (train.jl:33)."""
    g_lo, g_hi = shard_range(n_items // group, rank, world)
    return g_lo * group, g_hi * group


def all_reduce_counts(counts: np.ndarray, device=None) -> np.ndarray:
    """sum (K,4) int64 occurrence counts over ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    t = torch.from_numpy(np.ascontiguousarray(counts, np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def all_reduce_mean_(tensor):
    """in-place average over ranks (filter-gradient all-reduce of a training step)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor)
        tensor /= dist.get_world_size()
    return tensor
