# round-2 final single-GPU record: GPU tests, default bench line, reference arm, config-3 row with 64 groups per step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r02_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref rc $?"
timeout 600 python bench.py --workload config3 --c3-groups 64 --c3-steps 40 --no-cpu-baseline > gpurun_out/r02_config3_g64_n1.json 2> gpurun_out/r02_config3_g64_n1.err; echo "c3 rc $?"
tail -c 400 gpurun_out/r02_config3_g64_n1.json
