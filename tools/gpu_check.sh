#!/bin/bash
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --cf --surfaces 60 --batch 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cf_b64.json 2> gpurun_out/bench_cf_b64.err; tail -c 700 gpurun_out/bench_cf_b64.json
timeout 900 python bench.py --schedule reference --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_hybrid_b32.json 2> gpurun_out/bench_hybrid_b32.err; tail -c 700 gpurun_out/bench_hybrid_b32.json
timeout 900 python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --steps-per-stage 1000 > gpurun_out/bench_b16_literal.json 2> gpurun_out/bench_b16_literal.err; tail -c 700 gpurun_out/bench_b16_literal.json
timeout 600 python bench.py --batch 16 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16_T4.json 2> gpurun_out/bench_b16_T4.err; tail -c 700 gpurun_out/bench_b16_T4.json
