#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_denoisers.py tests/test_gpu_l4000.py tests/test_gpu_compaction.py tests/test_gpu_ops.py -x -q 2>&1 | tail -5
timeout 120 python tools/profile_forward.py --kind edgepos --batch 64 --iters 3 --time
timeout 120 python tools/profile_forward.py --kind surfpos --batch 64 --surfaces 30 --iters 50 --time
BG_ATTN_POLY=3 B=64 timeout 200 python tools/attn_check.py 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 5 -c 1 -o gpurun_out/r02_attn_b256 -f env B=256 python tools/attn_check.py > gpurun_out/r02_attn_ncu.out 2>&1
timeout 600 python bench.py --workload surfpos > gpurun_out/bench_surfpos2.json 2> gpurun_out/bench_surfpos2.err; tail -c 600 gpurun_out/bench_surfpos2.json
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
