#!/bin/bash
# N-GPU check of the bench contract (torchrun, NCCL): weak scaling + the final all_gather inside the e2e region
N=${N:-2}; B=${B:-64}
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --batch $B --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_b$B.json 2> gpurun_out/bench_${N}gpu_b$B.err
tail -c 1200 gpurun_out/bench_${N}gpu_b$B.json; tail -3 gpurun_out/bench_${N}gpu_b$B.err
