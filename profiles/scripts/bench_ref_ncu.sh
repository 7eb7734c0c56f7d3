set -x
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_tc$ -s 2 -c 1 -o gpurun_out/prof_scan_tc python bench.py --steps 1 --warmup 0 --no-cpu-baseline --workload scan --nseq 1000000 > gpurun_out/tc_ncu_full.log 2>&1
tail -3 gpurun_out/tc_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 20 --nseq 2000000 > gpurun_out/ncu_bench_tc.log 2>&1
