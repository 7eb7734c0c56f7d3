"""fp32 code retrieval throughput vs groups per launch (18 000 x 100 bp): python profiles/scripts/time_codes_groups.py [groups ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
N = 18000
a = synth.planted_gapped(N, 100, 2); seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(0))
for G in [int(x) for x in (sys.argv[1:] or ["100", "148", "250", "296", "500", "1000"])]:
    m = mb._lib.CscModel(ctx, hp, 100, n_groups=G, forward_only=True); m.set_params(cdl.flat)
    m.codes(seqs)
    t0 = time.perf_counter(); c = m.codes(seqs); dt = time.perf_counter() - t0
    print(f"groups per launch {G}: {N/dt:.0f} seq/s ({dt*1e3:.1f} ms), {len(c)} codes", flush=True)
    m.free()
