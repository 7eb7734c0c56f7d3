"""BASELINE config 2 end to end: synthetic FASTA 20 000 x 100 bp with a planted gapped motif, discover_motifs(fasta, outdir;
num_epochs=10) on one B200, wall time per stage."""
import os, sys, time, tempfile, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import synth, wrap, loadfasta, model, extract

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
a = synth.planted_gapped(N, 100, 2)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "reads.fa")
with open(fa, "w") as io:
    for i, row in enumerate(a):
        io.write(f">seq{i}\n{row.tobytes().decode()}\n")
stages = {}
def timed(mod, name, key):
    f = getattr(mod, name)
    def g(*args, **kw):
        t0 = time.perf_counter(); r = f(*args, **kw); stages[key] = stages.get(key, 0.0) + time.perf_counter() - t0; return r
    setattr(mod, name, g)
timed(loadfasta, "FASTA_DNA", "load_fasta"); timed(loadfasta, "get_data_bg", "background")
timed(model, "train_ucdl", "train_ucdl"); timed(model, "code_retrieval", "code_retrieval")
timed(extract, "run_thru", "run_thru_total"); timed(wrap, "render_result_", "render_result")
t0 = time.perf_counter()
ms, out = wrap.discover_motifs(fa, os.path.join(tmp, "out"), num_epochs=epochs, rng=np.random.default_rng(1))
total = time.perf_counter() - t0
stages["run_thru_without_code_retrieval"] = stages["run_thru_total"] - stages.get("code_retrieval", 0.0)
print(json.dumps({"config": f"{N} x 100 bp, num_epochs={epochs}", "total_s": round(total, 2), "stages_s": {k: round(v, 3) for k, v in stages.items()},
                  "motifs": ms.num_motifs, "significant": int(sum(p < 1e-5 for p in out["pvec"])),
                  "files": sorted(os.listdir(os.path.join(tmp, "out")))[:6]}))
