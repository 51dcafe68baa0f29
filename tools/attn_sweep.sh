#!/bin/bash
# one gpurun call: parity + timing of the attention op under several BG_ATTN_* settings (one process each: the knobs are
# read once per process), then the clock64 trace of the default kernel (debug build: make -C brepgen_b200/csrc trace)
#   gpurun --timeout 900 -- 'bash tools/attn_sweep.sh > gpurun_out/attn_sweep.log 2>&1'
run() { echo "=== $*"; env "$@" timeout 240 python tools/attn_check.py 2>&1 | tail -9; }
run B=64
run B=64 BG_ATTN_POLY=0
run B=64 BG_ATTN_PP=0
run B=256
[ -f brepgen_b200/libbrepgen_trace.so ] && timeout 240 python tools/attn_trace.py 2>&1 | tail -30
