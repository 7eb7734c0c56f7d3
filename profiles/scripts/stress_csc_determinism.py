"""Race hunt: the same step (same parameters, same batch) repeated many times must return the same loss bits (the fused forward is
deterministic) and gradients equal up to the float atomics left in the tape's DF / warm-up adjoints.
usage: python profiles/scripts/stress_csc_determinism.py [Lb] [repeats]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth

Lb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(600, Lb, 2)
seqs = ctx.seqs_from_ascii(a)
m = mb._lib.CscModel(ctx, hp, Lb)
rng = np.random.default_rng(0)
bad = 0
for trial in range(6):
    cdl = mdl.ucdl(hp, np.random.default_rng(trial))
    m.set_params(cdl.flat)
    for _ in range(30 * trial):                      # move away from the initialisation
        m.step_begin(seqs, rng.permutation(600)[:6]); m.adabelief_step()
    idx = rng.permutation(600)[:6]
    l0, g0 = m.loss_grad(seqs, idx)
    x0 = m.get_buffer("x", 6 * (Lb - 18) * hp.K)
    worst = 0.0
    for r in range(reps):
        l1, g1 = m.loss_grad(seqs, idx)
        x1 = m.get_buffer("x", 6 * (Lb - 18) * hp.K)
        if not np.array_equal(l0, l1) or not np.array_equal(x0, x1):
            bad += 1
            print(f"trial {trial} rep {r}: forward differs: loss {l0.ravel()} vs {l1.ravel()}, x entries differing {(x0 != x1).sum()}")
        worst = max(worst, float(np.abs(g1 - g0).max() / np.abs(g0).max()))
    print(f"trial {trial}: {reps} repeats, forward mismatches so far {bad}, worst relative gradient deviation {worst:.2e}")
print("RESULT", "ok" if bad == 0 else f"{bad} forward mismatches")
