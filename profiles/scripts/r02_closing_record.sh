# round-2 closing record on one B200 (final library): smoke, GPU tests, default bench line, reference arm, launch list of a 64-group step, ncu full of k_corr2d_s
# usage (GPU box, repo root): bash profiles/scripts/r02_closing_record.sh
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r02_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref rc $?"
bash profiles/scripts/r02_groups_launches.sh 200 64 > gpurun_out/r02_groups_summary.txt 2>/dev/null; head -3 gpurun_out/r02_groups_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_corr2d_s -s 3 -c 1 -f -o gpurun_out/r02_corr2d_s_full python profiles/scripts/prof_csc_groups.py 200 64 > gpurun_out/ncu_c2s.log 2>&1; echo "ncu rc $?"
