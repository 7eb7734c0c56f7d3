"""training step time vs groups per step (each group = one reference batch of 6; gradients averaged over groups)."""
import numpy as np, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
Lb = 100
a = synth.planted_gapped(12000, Lb, 2); seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(0))
for G in [int(x) for x in (sys.argv[1:] or ["1", "8", "16", "32", "64", "128"])]:
    m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G); m.set_params(cdl.flat)
    rng = np.random.default_rng(1)
    for it in range(5):
        m.step_begin(seqs, rng.permutation(12000)[:6 * G]); m.adabelief_step()
    n = 60; t0 = time.perf_counter()
    for it in range(n):
        m.step_begin(seqs, rng.permutation(12000)[:6 * G]); loss, l1 = m.adabelief_step()
    dt = (time.perf_counter() - t0) / n
    print(f"G={G}: {dt*1e3:.3f} ms/step, {6*G/dt:.0f} seq/s, loss {loss:.3f}", flush=True)
    m.free()
