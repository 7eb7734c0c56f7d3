set -x
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 20 > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 20 > gpurun_out/ncu_bench.log 2>&1
tail -c 600 gpurun_out/bench_n1.json
