"""triplet enumeration on the device: 20 000 sequences x 32 random code components (10^8 triplets at the lowest quantile),
the eight quantile passes of run_thru (_g1_obtain_coutmats.jl:131-160)."""
import numpy as np, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import motifs_jl_b200 as mb
from motifs_jl_b200 import extract
from motifs_jl_b200._lib import CODE_DTYPE
ctx = mb.Context(0)
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000; per = 32
n = N * per
codes = np.zeros(n, CODE_DTYPE)
codes["seq"] = np.repeat(np.arange(N), per)
codes["fil"] = np.sort(rng.integers(0, 24, (N, per)), axis=1).ravel()
codes["position"] = rng.integers(0, 82, n)
codes["mag_f16"] = rng.random(n).astype(np.float16).view(np.uint16)
# a planted word in a third of the sequences
for s in range(0, N, 3):
    codes["fil"][s * per: s * per + 3] = (2, 2, 2); codes["position"][s * per: s * per + 3] = (10, 17, 29)
    codes["mag_f16"][s * per: s * per + 3] = np.float16(2.0).view(np.uint16)
t0 = time.perf_counter()
for q in (0.75, 0.65, 0.5, 0.45, 0.35, 0.25, 0.15, 0.05):
    t1 = time.perf_counter()
    cf = extract.filter_code_components_using_quantile(codes, q)
    t2 = time.perf_counter()
    t = extract.enumerate_triplets_gpu(ctx, cf)
    t3 = time.perf_counter()
    ek = extract.get_enriched_keys_gpu(t, max_word_combinations=1000)
    t4 = time.perf_counter()
    vals = t.values(ek["key"], total=int(ek["count"].sum())) if len(ek) else []
    t5 = time.perf_counter()
    print(f"q={q}: {len(cf)} codes, {t.n_triplets} triplets, {len(ek)} enriched keys | filter {t2-t1:.3f}s count {t3-t2:.3f}s frequent {t4-t3:.3f}s values {t5-t4:.3f}s", flush=True)
    t.free()
print(f"total {time.perf_counter() - t0:.2f} s")
