"""A few many-group CSC training steps for ncu (graph nodes profiled): python profiles/scripts/prof_csc_groups.py [Lb] [groups]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
Lb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(6 * G + 100, Lb, 2)
seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(2))
m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G)
m.set_params(cdl.flat)
rng = np.random.default_rng(0)
for _ in range(3):
    m.step_begin(seqs, rng.permutation(6 * G + 100)[:6 * G]); print(m.adabelief_step())
