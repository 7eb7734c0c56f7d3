# same as run_roles.sh but with the in-tree library built by `make TCS_PROFILE=1 -B` (rebuild with plain `make -B` afterwards)
export MB200_SCAN_TC_STATS=1 MB200_SCAN_TC_DEBUG=2
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --workload scan --nseq 1000000 > gpurun_out/tc_dbg.json 2> gpurun_out/tc_dbg.err
grep tcdbg gpurun_out/tc_dbg.err | head -148 | awk 'NR%12==1'
grep "tensor-core" gpurun_out/tc_dbg.err | tail -2
