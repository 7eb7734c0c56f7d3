export MB200_SCAN_TC_STATS=1
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload scan --nseq 2000000 > gpurun_out/tc_bench_small.json 2> gpurun_out/tc_bench_small.err
tail -c 400 gpurun_out/tc_bench_small.json; tail -3 gpurun_out/tc_bench_small.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tc_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload scan --nseq 2000000 > gpurun_out/tc_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/tc_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    u=r[ui]; v = v/1e3 if u=='ns' else (v if u=='us' else v*1e3 if u=='ms' else v)
    agg[r[ki][:60]][0]+=1; agg[r[ki][:60]][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:60s} {n:5d} {t/1e3:10.3f} ms")
PY
