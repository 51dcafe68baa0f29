#!/bin/bash
# one gpurun call: the whole GPU test suite, smoke(), the default bench (BASELINE configs[2], B = 256) and the reference arm
#   gpurun --timeout 3000 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1'
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3200 gpurun_out/bench_default.json
