import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(3000, 100, 2); seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(0))
m = mb._lib.CscModel(ctx, hp, 100, n_groups=500, forward_only=True, tensor_cores=(len(sys.argv) > 1 and sys.argv[1] == "tc")); m.set_params(cdl.flat)
m.codes(seqs)
