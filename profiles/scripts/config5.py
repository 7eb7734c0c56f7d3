"""BASELINE config 5: 250 Mbp chromosome-length synthetic sequence, K=50 PWMs (len 8-40), tiled scan with motif-length halos,
per-motif hit / unique / coverage counts on the sequence and on its 1-mer shuffle, Fisher enrichment.
Run on the GPU box from the repo root:  python profiles/scripts/config5.py [Lb]"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb  # noqa: E402
from motifs_jl_b200 import inference, synth  # noqa: E402
from oracle import scan_oracle as so  # noqa: E402

Lb = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
ctx = mb.Context(0)
rng = np.random.default_rng(5)
a = synth.random_ascii(1, Lb, 5)
fams = [b"TGACGTCATTGACGTCA", b"GGGACTTTCC", b"CACGTGACCGGAAGT"]
for f in fams:
    site = np.frombuffer(f, np.uint8)
    for p in rng.integers(0, Lb - 40, Lb // 50_000):
        a[0, p:p + len(site)] = site
bg = rng.permutation(a[0])[None, :]
cms = [synth.count_matrix_from_sites([f.decode()] * 30) for f in fams] + synth.random_count_matrices(47, 8, 40, 6)
ms = synth.motifs_from_count_matrices(cms)
pw, lens = so.pack_pwms(ms.pwms)
thr = synth.stated_thresholds(ms, 0.7)
out = {"Lb": Lb, "K": len(lens)}
res = []
for name, rows in (("fg", a), ("bg", bg)):
    t0 = time.perf_counter()
    seqs = ctx.seqs_from_ascii(rows)
    t_up = time.perf_counter() - t0
    ctx.scan(seqs, pw, lens, thr, want_hits=False)
    t0 = time.perf_counter()
    _, c = ctx.scan(seqs, pw, lens, thr, want_hits=False)
    dt = time.perf_counter() - t0
    tm, ln = ctx.last_timing()
    out[name] = {"upload_s": t_up, "scan_s": dt, "bp_per_s": Lb / dt, "kernel_ms": tm, "hits_planted": c[:3, 0].tolist()}
    res.append(c)
    sub = rows[:, :2_000_000]                                  # parity on the first 2 Mbp against the oracle
    s2 = ctx.seqs_from_ascii(sub)
    _, c2 = ctx.scan(s2, pw, lens, thr, want_hits=False)
    _, oc = so.scan(pw, lens, so.ascii_to_codes(sub), thr, want_hits=False)
    out[name]["first_2Mbp_counts_match_oracle"] = bool(np.array_equal(c2, oc))
    s2.free()
    seqs.free()
p = inference.fisher_pvec(res[0][:, 2], res[1][:, 2], SimpleNamespace(N=1, L=Lb, N_test=0))
out["fisher_p_planted"] = p[:3].tolist()
out["n_significant_1e-5"] = int((p < 1e-5).sum())
print(json.dumps(out))
