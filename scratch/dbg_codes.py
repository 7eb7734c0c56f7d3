import numpy as np, sys
sys.path.insert(0, '/root/repo')
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
from oracle import csc_oracle as co, scan_oracle as so
from types import SimpleNamespace
ctx = mb.Context(0)
hp = mdl.Hyperparam(); ohp = co.Hyperparam()
a = synth.planted_gapped(50, 100, 7); seqs = ctx.seqs_from_ascii(a); flat = co.init_params(ohp, 7)
cdl = mdl.ucdl(hp, np.random.default_rng(0)); cdl.flat[:] = flat
data = SimpleNamespace(N=50, L=100, seqs=seqs)
for gpc in (1, 3, 8):
    got = mdl.code_retrieval(data, cdl, hp, groups_per_call=gpc)
    exp = co.code_retrieval(so.ascii_to_codes(a), flat, ohp)
    gm, em = got["mag_f16"].view(np.float16).astype(np.float32), exp["mag_f16"].view(np.float16).astype(np.float32)
    d = np.abs(gm-em)
    print(gpc, len(got), len(exp), "maxdiff", d.max(), "n>2e-6", (d>2e-6).sum(), "seqs with big diff", np.unique(got["seq"][d>2e-6])[:20])
    i = d.argmax(); print(got[i], exp[i], gm[i], em[i])
