set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --train-steps 300 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 600 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 --train-steps 300 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --train-steps 300 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
