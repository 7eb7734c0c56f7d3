# round-2 state check on the GPU box: GPU tests, default bench line, fused CSC step timing, ncu launch list of the fused step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "tests rc $?"
timeout 900 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc $?"
timeout 300 python profiles/scripts/time_csc_fused.py 100 > gpurun_out/r02b_t100.txt 2>&1
timeout 300 python profiles/scripts/time_csc_fused.py 200 > gpurun_out/r02b_t200.txt 2>&1
timeout 200 python profiles/scripts/prof_csc_fused.py > gpurun_out/plain_csc.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 400 --csv --log-file gpurun_out/r02b_fused_launches.csv python profiles/scripts/prof_csc_fused.py > gpurun_out/ncu_csc.log 2>&1
tail -3 gpurun_out/r02b_tests.log; cat gpurun_out/r02b_t100.txt gpurun_out/r02b_t200.txt
