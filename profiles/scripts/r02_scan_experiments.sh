for lib in "" build_hk/lib_halfk.so build_hd/lib_halfdrain.so build_nd/lib_nodrain.so; do
  MB200_LIBRARY=$lib timeout 600 python bench.py --workload scan --nseq 3000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_scan.json 2> gpurun_out/exp_scan.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/exp_scan.json").read().strip().splitlines()[-1])
print("${lib:-production}", "ms_per_step", round(d["ms_per_step"],2), "Gbp/s", round(d["value"]/1e9,3), "k_scan_tc ms/launch", round(d["roofline"]["ms_per_launch"],3))
PY
done
