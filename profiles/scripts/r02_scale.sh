# usage (GPU box with N GPUs): bash profiles/scripts/r02_scale.sh N   -- the driver's command at N ranks: bench line and reference arm
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc $?"
run 29512 --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; echo "ref rc $?"
tail -c 300 gpurun_out/r02_bench_n$N.json; tail -2 gpurun_out/r02_bench_n$N.err
