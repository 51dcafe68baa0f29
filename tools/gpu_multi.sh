#!/bin/bash
# 2-GPU check of the bench contract (torchrun, NCCL): weak scaling + the final all_gather inside the e2e region
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --batch 64 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_b64.json 2> gpurun_out/bench_2gpu_b64.err
tail -c 1500 gpurun_out/bench_2gpu_b64.json; tail -5 gpurun_out/bench_2gpu_b64.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
tail -c 600 gpurun_out/bench_2gpu_ref.json
