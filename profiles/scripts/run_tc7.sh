timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print("value",d['value'],"ms/step",d['ms_per_step'],"e2e",d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
r=d['roofline']; print({k:r.get(k) for k in ('achieved','frac','ms_per_launch','kernel_share_of_step','verify_kernel_share_of_step','count_kernel_share_of_step')})
print(d['checks'], d['training']['value'])
PY
