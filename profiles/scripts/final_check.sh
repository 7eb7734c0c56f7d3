set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 20 --nseq 2000000 > gpurun_out/ncu_bench_tc.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print("value",d['value'],"ms/step",d['ms_per_step'],"e2e",d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
r=d['roofline']; print({k:r.get(k) for k in ('achieved','frac','ms_per_launch','kernel_share_of_step','traffic')})
print(d['checks'], d['training']['value'], d['training']['e2e']['value'])
PY
