# per-role clocks of k_scan_tc: needs a TCS_PROFILE build beside the production library:
#   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DTCS_PROFILE=1 -shared -o profiles/scripts/_prof_libmotifs_b200.so motifs.jl_b200/csrc/*.cu
export MB200_LIBRARY=$PWD/profiles/scripts/_prof_libmotifs_b200.so MB200_SCAN_TC_STATS=1 MB200_SCAN_TC_DEBUG=2
timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --workload scan --nseq 1000000 > gpurun_out/tc_dbg.json 2> gpurun_out/tc_dbg.err
grep -c tcdbg gpurun_out/tc_dbg.err; grep "tensor-core" gpurun_out/tc_dbg.err | tail -1
