export MB200_SCAN_TC_STATS=1
for mode in 0 1 2 0; do
MB200_SCAN_TC_DRAIN=$mode timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload scan --nseq 2000000 2>&1 >/dev/null | grep "tensor-core" | tail -1
done
MB200_SCAN_TC_DRAIN=1 MB200_SCAN_TC_MMAREP=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload scan --nseq 2000000 2>&1 >/dev/null | grep "tensor-core" | tail -1
