import numpy as np, sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
for Lb, G in ((100, 1), (100, 4), (100, 16), (200, 1)):
    a = synth.planted_gapped(6000, Lb, 2); seqs = ctx.seqs_from_ascii(a)
    cdl = mdl.ucdl(hp, np.random.default_rng(0))
    m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G); m.set_params(cdl.flat)
    rng = np.random.default_rng(1)
    for it in range(5):
        m.step_begin(seqs, rng.permutation(6000)[:6*G]); m.adabelief_step()
    t0 = time.perf_counter(); n = 200
    for it in range(n):
        m.step_begin(seqs, rng.permutation(6000)[:6*G]); loss, l1 = m.adabelief_step()
    dt = (time.perf_counter() - t0) / n
    print(f"Lb={Lb} G={G}: {dt*1e3:.3f} ms/step, {6*G/dt:.0f} seq/s, loss {loss:.3f} l1 {l1:.2f}")
    # forward-only code retrieval throughput
    mf = mb._lib.CscModel(ctx, hp, Lb, n_groups=min(500, 6000//6), forward_only=True); mf.set_params(cdl.flat)
    mf.codes(seqs)
    t0 = time.perf_counter(); codes = mf.codes(seqs); dt = time.perf_counter() - t0
    print(f"   code retrieval: {6000/dt:.0f} seq/s, {len(codes)} codes")
    m.free(); mf.free(); seqs.free()
