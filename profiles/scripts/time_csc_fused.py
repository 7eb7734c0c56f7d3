"""Times the CSC training step (batch 6, resident sequences) with the fused persistent kernels and with the kernel-per-op tape.
usage (GPU box, repo root): python profiles/scripts/time_csc_fused.py [Lb] [groups]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth

Lb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(2000, Lb, 2)
seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(2))
for fused in (True, False):
    m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G, fused=fused)
    m.set_params(cdl.flat)
    rng = np.random.default_rng(0)
    for _ in range(50):
        m.step_begin(seqs, rng.permutation(2000)[:6 * G]); m.adabelief_step()
    l0 = ctx.last_timing()[1]["csc"]
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        m.step_begin(seqs, rng.permutation(2000)[:6 * G]); loss, l1 = m.adabelief_step()
    dt = time.perf_counter() - t0
    print(f"Lb={Lb} groups={G} fused={fused}: {1e3 * dt / n:.4f} ms per optimiser step, {(ctx.last_timing()[1]['csc'] - l0) / n:.1f} kernels per step, loss {loss:.4f}")
    m.free()
