# per-phase clocks of the three fused CSC kernels (FZ_PROFILE build through MB200_LIBRARY) and the ncu launch list of the production step
mkdir -p gpurun_out
MB200_LIBRARY=build_fz/libmotifs_b200_fzprof.so timeout 200 python profiles/scripts/prof_csc_fused.py > gpurun_out/r02c_fzprof.txt 2>&1
timeout 200 python profiles/scripts/prof_csc_fused.py > gpurun_out/plain_csc.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 400 --csv --log-file gpurun_out/r02c_fused_launches.csv python profiles/scripts/prof_csc_fused.py > gpurun_out/ncu_csc.log 2>&1
grep -E "fzd|fz\] total|fzb\] total" gpurun_out/r02c_fzprof.txt | tail -16
