#!/bin/bash
set -x
timeout 200 python tools/gemm_time.py
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_denoisers.py -x -q 2>&1 | tail -3
timeout 120 python tools/profile_forward.py --kind edgepos --batch 64 --iters 3 --time
