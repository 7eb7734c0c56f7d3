timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_scan_tc$ -s 1 -c 1 -o gpurun_out/prof_scan_tc_final python bench.py --steps 1 --warmup 0 --no-cpu-baseline --workload scan --nseq 600000 > gpurun_out/tc_ncu_full.log 2>&1
tail -2 gpurun_out/tc_ncu_full.log
