"""world_size-2 gloo test of the N>1 host logic: sequence sharding + one count reduction reproduces the
single-rank result (counts are sums over sequences).  The per-shard compute is the CPU oracle here; on GPUs the
same code path calls the library (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from motifs_jl_b200 import parallel, synth
    from oracle import scan_oracle as so
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = synth.random_ascii(101, 80, 5)
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(6, 8, 20, 6))
    pw, lens = so.pack_pwms(ms.pwms)
    lo, hi = parallel.shard_range(len(a), rank, world)
    _, c = so.scan(pw, lens, so.ascii_to_codes(a[lo:hi]), want_hits=False)
    tot = parallel.all_reduce_counts(c)
    if rank == 0:
        _, full = so.scan(pw, lens, so.ascii_to_codes(a), want_hits=False)
        q.put(bool(np.array_equal(tot, full)))
    dist.destroy_process_group()


def test_sharded_counts_sum_to_full_counts():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_shard_helpers():
    from motifs_jl_b200 import parallel
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            cuts = [parallel.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
    assert parallel.shard_groups(20, 6, 0, 2) == (0, 6) and parallel.shard_groups(20, 6, 1, 2) == (6, 18)


def _grad_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from motifs_jl_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(30433, dtype=torch.float32) * (rank + 1)
    parallel.all_reduce_mean_(g)
    if rank == 0:
        q.put(bool(torch.allclose(g, torch.arange(30433, dtype=torch.float32) * 1.5)))
    dist.destroy_process_group()


def test_gradient_allreduce_is_a_mean():
    """the per-step filter-gradient all-reduce averages the ranks' gradient vectors."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
