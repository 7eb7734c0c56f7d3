export MB200_SCAN_TC_STATS=1
MB200_SCAN_TC_ACCW=128 timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
for w in 256 0 128 256 0; do
echo "ACCW=$w"; MB200_SCAN_TC_ACCW=$w timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload scan --nseq 2000000 2>&1 >/dev/null | grep "tensor-core" | tail -1
done
