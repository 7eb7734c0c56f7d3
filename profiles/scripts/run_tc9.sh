export MB200_SCAN_TC_STATS=1
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
for pr in 0 1 0 1; do
echo "PAIR=$pr"; MB200_SCAN_TC_PAIR=$pr timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload scan --nseq 2000000 2>&1 >/dev/null | grep "tensor-core" | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tc2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload scan --nseq 2000000 > /dev/null 2>&1
grep -c "k_scan_tc2" gpurun_out/tc2_launches.csv
