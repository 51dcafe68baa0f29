#!/bin/bash
# one gpurun call: the GPU test suite, the BASELINE configs[1] line (eager vs graph) and a short cascade bench
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1'
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --workload surfpos > gpurun_out/bench_surfpos.json 2> gpurun_out/bench_surfpos.err; tail -c 1500 gpurun_out/bench_surfpos.json
timeout 900 python bench.py --batch 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; tail -c 3000 gpurun_out/bench_b64.json
