import sys
import numpy as np
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(2000, 100, 2)
seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(2))
m = mb._lib.CscModel(ctx, hp, 100)
m.set_params(cdl.flat)
rng = np.random.default_rng(0)
for it in range(int(sys.argv[1])):
    m.step_begin(seqs, rng.permutation(2000)[:6]); m.adabelief_step()
