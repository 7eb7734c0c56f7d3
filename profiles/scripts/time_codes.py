"""code retrieval throughput: fp32 SIMT path vs tcgen05 BF16 path (forward_only=2), 6000 x 100 bp, 500 groups per call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(6000, 100, 2); seqs = ctx.seqs_from_ascii(a)
cdl = mdl.ucdl(hp, np.random.default_rng(0))
for tc in (False, True):
    m = mb._lib.CscModel(ctx, hp, 100, n_groups=500, forward_only=True, tensor_cores=tc); m.set_params(cdl.flat)
    m.codes(seqs)
    t0 = time.perf_counter(); c = m.codes(seqs); dt = time.perf_counter() - t0
    tm, ln = ctx.last_timing()
    print(f"tensor_cores={tc}: {6000/dt:.0f} seq/s ({dt*1e3:.1f} ms, device {tm['csc']:.1f} ms), {len(c)} codes")
    m.free()
