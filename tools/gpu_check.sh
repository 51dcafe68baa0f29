#!/bin/bash
# one gpurun call: attention parity + timing (new kernel vs the round-1 kernel), GEMM timing, then the GPU test suite
#   gpurun --timeout 1200 -- 'bash tools/gpu_check.sh > gpurun_out/gpu_check.log 2>&1'
set -x
timeout 300 python tools/attn_check.py
BG_ATTN_V=5 timeout 300 python tools/attn_check.py
BG_ATTN_POLY=0 timeout 300 python tools/attn_check.py
BG_ATTN_POLY=2 timeout 300 python tools/attn_check.py
B=256 timeout 300 python tools/attn_check.py
timeout 300 python tools/gemm_time.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
