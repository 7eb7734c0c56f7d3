mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc $?"; tail -1 gpurun_out/r02_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r02_bench_n1.err
