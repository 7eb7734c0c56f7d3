"""H2D probe: pinned 2 GB ASCII -> device, raw torch copy vs the library's chunked asynchronous upload + pack."""
import time, torch, numpy as np
import motifs_jl_b200 as mb
ctx = mb.Context(0)
N, Lb = 10_000_000, 200
host = torch.empty((N, Lb), dtype=torch.uint8, pin_memory=True)
host.random_(65, 66)
host[:] = 65
dev = torch.empty((N, Lb), dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("torch pinned H2D 2 GB: %.1f ms  %.1f GB/s" % (dt * 1e3, N * Lb / dt / 1e9))
for _ in range(3):
    t0 = time.perf_counter(); s = ctx.seqs_from_host_ptr(host.data_ptr(), N, Lb, wait=True); dt = time.perf_counter() - t0
    print("library upload+pack (wait): %.1f ms  %.1f GB/s" % (dt * 1e3, N * Lb / dt / 1e9)); s.free()
