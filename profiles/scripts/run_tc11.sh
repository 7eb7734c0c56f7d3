export MB200_SCAN_TC_STATS=1
MB200_SCAN_TC_BSWZ=1 timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -4
for z in 0 1 0 1; do
echo "BSWZ=$z"; MB200_SCAN_TC_BSWZ=$z timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload scan --nseq 2000000 2>gpurun_out/z.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"; grep "tensor-core" gpurun_out/z.err | tail -1
done
