"""Multi-GPU plumbing (one process per GPU).  The data-path collectives live INSIDE libmotifs_b200 (csrc/comm.cu, one NCCL
communicator per ctx): the scan shards sequences over ranks with no collective and sums the per-motif counts once per call
(MB200_SCAN_REDUCE), training averages the filter gradients with one all-reduce per step inside mb200_csc_adabelief_step,
code retrieval shards whole batches (mb200_csc_codes_sharded) — SURVEY §8e.  What is left here is the host-side bootstrap
(handing rank 0's 128-byte communicator id to the other processes) and the shard arithmetic; torch.distributed is used only
as that bootstrap channel when it happens to be initialised (any backend, gloo in the CPU tests)."""
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


def init_comm(ctx, rank: int | None = None, world: int | None = None, unique_id: bytes | None = None):
    """create ctx's communicator.  Bootstrap of the id, in this order: the `unique_id` argument (rank 0 made it with
    Context.comm_unique_id() and the host shipped it); torch.distributed when initialised; else a TCPStore at
    MASTER_ADDR:MASTER_PORT+1 with RANK / WORLD_SIZE from the environment (what torchrun exports).  Returns (rank, world)."""
    import os
    if unique_id is None:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
            if world == 1:
                return 0, 1
            dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            buf = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).clone()
            buf = buf.to(dev)
            dist.broadcast(buf, 0)
            unique_id = bytes(buf.cpu().numpy().tobytes())
        else:
            rank = int(os.environ.get("RANK", "0")) if rank is None else rank
            world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
            if world == 1:
                return 0, 1
            from datetime import timedelta
            store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1, world,
                                  is_master=(rank == 0), timeout=timedelta(seconds=120))
            if rank == 0:
                store.set("mb200_comm_id", ctx.comm_unique_id())
            unique_id = bytes(store.get("mb200_comm_id"))
    if rank is None or world is None:
        raise ValueError("rank and world are needed with an explicit unique_id")
    ctx.comm_init(unique_id, rank, world)
    return rank, world


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
