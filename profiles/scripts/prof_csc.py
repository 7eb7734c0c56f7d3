import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(600, 100, 2); seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(0))
m = mb._lib.CscModel(ctx, hp, 100, n_groups=1); m.set_params(cdl.flat)
for it in range(3):
    m.step_begin(seqs, np.arange(6) + 6 * it); m.adabelief_step()
