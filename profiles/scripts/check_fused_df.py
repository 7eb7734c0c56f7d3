"""Gradient of one CSC step with the fused DF reverse kernel against the tape's DF reverse pass (both with the fused forward + XYZ reverse
kernels), per parameter block, and against the oracle; then step timing.  usage: python profiles/scripts/check_fused_df.py [Lb] [groups]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import motifs_jl_b200 as mb
from motifs_jl_b200 import model as mdl, synth
from oracle import csc_oracle as co, scan_oracle as so

Lb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = mb.Context(0)
hp = mdl.Hyperparam()
a = synth.planted_gapped(60, Lb, 3)
seqs = ctx.seqs_from_ascii(a)
ohp = co.Hyperparam(**{k: getattr(hp, k) for k in ("filter_len", "M", "h", "K", "q", "batch_size", "num_pass_xyz", "num_pass_df", "magnifying_factor", "gamma")})
flat = co.init_params(ohp, 3)
idx = np.random.default_rng(1).permutation(60)[:6 * G]
res = {}
for name, kw in (("fused_df", {}), ("tape_df", {"fused_df": False}), ("tape", {"fused": False})):
    m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G, **kw)
    m.set_params(flat)
    res[name] = m.loss_grad(seqs, idx)
    m.free()
og = None
if G == 1:
    oloss, og, aux = co.loss_and_grad(so.ascii_to_codes(a)[idx], flat, ohp)
    print("loss", res["fused_df"][0][0], "oracle", oloss)
o = 0
for name, n in co.param_sizes(ohp).items():
    ref = res["tape_df"][1][o:o + n]; new = res["fused_df"][1][o:o + n]; tp = res["tape"][1][o:o + n]
    sc = max(np.abs(ref).max(), 1e-12)
    line = f"{name:8s} n={n:6d} |ref|max {sc:.3e}  fused_df-tape_df {np.abs(new - ref).max() / sc:.2e}  tape_df-tape {np.abs(ref - tp).max() / sc:.2e}"
    if og is not None:
        osc = max(np.abs(og[o:o + n]).max(), 1e-12)
        line += f"  fused_df-oracle {np.abs(new - og[o:o + n]).max() / osc:.2e}  tape-oracle {np.abs(tp - og[o:o + n]).max() / osc:.2e}"
    print(line)
    o += n
cdl = mdl.ucdl(hp, np.random.default_rng(2))
b = synth.planted_gapped(2000, Lb, 2)
seqs2 = ctx.seqs_from_ascii(b)
for name, kw in (("fused_df", {}), ("tape_df", {"fused_df": False})):
    m = mb._lib.CscModel(ctx, hp, Lb, n_groups=G, **kw)
    m.set_params(cdl.flat)
    rng = np.random.default_rng(0)
    for _ in range(50):
        m.step_begin(seqs2, rng.permutation(2000)[:6 * G]); m.adabelief_step()
    l0 = ctx.last_timing()[1]["csc"]
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        m.step_begin(seqs2, rng.permutation(2000)[:6 * G]); loss, l1 = m.adabelief_step()
    dt = time.perf_counter() - t0
    print(f"Lb={Lb} groups={G} {name}: {1e3 * dt / n:.4f} ms per optimiser step, {(ctx.last_timing()[1]['csc'] - l0) / n:.1f} kernels per step, loss {loss:.4f}")
    m.free()
