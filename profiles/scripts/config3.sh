set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --workload train --train-nseq 200000 --train-seqlen 200 --train-steps 300 --no-cpu-baseline > gpurun_out/config3_n8.json 2> gpurun_out/config3_n8.err
tail -c 1500 gpurun_out/config3_n8.json; tail -3 gpurun_out/config3_n8.err
python bench.py --steps 3 --warmup 3 --workload train --train-nseq 200000 --train-seqlen 200 --train-steps 300 --no-cpu-baseline > gpurun_out/config3_n1.json 2> gpurun_out/config3_n1.err
tail -c 1500 gpurun_out/config3_n1.json
