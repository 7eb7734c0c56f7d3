# A/B of two builds of the library on ONE box (box-to-box variation is ~15 %): build the other version with
#   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o profiles/scripts/_old_libmotifs_b200.so <its csrc>/*.cu
# (git-ignored, travels with gpurun) and run this script from the repo root.
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -2
for lib in old new old new; do
if [ $lib = old ]; then export MB200_LIBRARY=$PWD/profiles/scripts/_old_libmotifs_b200.so; else unset MB200_LIBRARY; fi
echo "LIB=$lib"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload scan --nseq 2000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'])"
done
