"""cProfile of discover_motifs on BASELINE config 2 (host-side hot spots around the GPU calls)."""
import os, sys, tempfile, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from motifs_jl_b200 import synth, wrap
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
a = synth.planted_gapped(N, 100, 2)
tmp = tempfile.mkdtemp(); fa = os.path.join(tmp, "reads.fa")
with open(fa, "w") as io:
    for i, row in enumerate(a):
        io.write(f">seq{i}\n{row.tobytes().decode()}\n")
pr = cProfile.Profile(); pr.enable()
wrap.discover_motifs(fa, os.path.join(tmp, "out"), num_epochs=10, rng=np.random.default_rng(1))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
