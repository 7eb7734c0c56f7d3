# Upper bound of what an FP8 (kind::f8f6f4) pre-filter could gain: the same scan with HALF the MMAs per accumulator use
# (make BUILD=build_hk LIB=build_hk/lib_halfk.so EXTRA_DEFS="-DTCS_EMULATE_HALF_K=1" build_hk/lib_halfk.so; the results of that build are wrong, only its timing is of interest)
mkdir -p gpurun_out
for lib in "" build_hk/lib_halfk.so; do
  MB200_LIBRARY=$lib MB200_SCAN_TC_STATS=1 timeout 600 python bench.py --workload scan --nseq 3000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/halfk_${lib:+hk}.json 2> gpurun_out/halfk_${lib:+hk}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/halfk_${lib:+hk}.json").read().strip().splitlines()[-1])
print("${lib:-production}", "ms_per_step", round(d["ms_per_step"],2), "Gbp/s", round(d["value"]/1e9,3), "k_scan_tc ms/launch", round(d["roofline"]["ms_per_launch"],3))
PY
done
