"""Where does the end-to-end scan step spend its time: upload (async) + scan, per-slot timers of the library."""
import time, torch, numpy as np, sys
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
import bench
from oracle import scan_oracle as so
ctx = mb.Context(0)
N, Lb = 4_000_000, 200
ms, thr = bench.make_motifs(500)
pw, lens = so.pack_pwms(ms.pwms)
host = torch.empty((N, Lb), dtype=torch.uint8, pin_memory=True)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8)
host.copy_(lut[torch.randint(0, 4, (N, Lb))])
for mode in ("wait", "async", "async", "resident"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = ctx.seqs_from_host_ptr(host.data_ptr(), N, Lb, wait=(mode != "async"))
    t1 = time.perf_counter()
    _, c = ctx.scan(s, pw, lens, thr, want_hits=False)
    t2 = time.perf_counter()
    if mode == "resident":
        _, c = ctx.scan(s, pw, lens, thr, want_hits=False); t2b = time.perf_counter(); print("  second resident scan %.1f ms" % ((t2b - t2) * 1e3))
    t, l = ctx.last_timing()
    s.free(); t3 = time.perf_counter()
    print(mode, "upload call %.1f ms, scan call %.1f ms, free %.1f ms | lib timers" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), {k: round(v, 1) for k, v in t.items()})
