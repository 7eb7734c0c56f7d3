# round-2 closing record on one B200: smoke, GPU tests, default bench line, reference arm, then ONE `ncu --set full` capture of the fused forward kernel
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r02_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref rc $?"
timeout 200 python profiles/scripts/prof_csc_fused.py > gpurun_out/plain_csc.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_csc_fused_fwd -s 2 -c 1 -f -o gpurun_out/r02_csc_fused_fwd_full python profiles/scripts/prof_csc_fused.py > gpurun_out/ncu_csc_full.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/ncu_csc_full.log
