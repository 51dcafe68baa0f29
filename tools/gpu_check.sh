#!/bin/bash
# one gpurun call: GPU test suite, then compaction / small-M timings
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 120 python tools/profile_forward.py --kind surfpos --batch 64 --surfaces 30 --iters 50 --time
BG_GEMM_SMALLM=0 timeout 120 python tools/profile_forward.py --kind surfpos --batch 64 --surfaces 30 --iters 50 --time
timeout 600 python bench.py --batch 32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --masks ragged --compact 1 > gpurun_out/bench_ragged_c1.json 2> gpurun_out/bench_ragged_c1.err; tail -c 1200 gpurun_out/bench_ragged_c1.json
timeout 600 python bench.py --batch 32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --masks ragged --compact 0 > gpurun_out/bench_ragged_c0.json 2> gpurun_out/bench_ragged_c0.err; tail -c 1200 gpurun_out/bench_ragged_c0.json
timeout 600 python bench.py --batch 32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_dense_b32.json 2> gpurun_out/bench_dense_b32.err; tail -c 1200 gpurun_out/bench_dense_b32.json
