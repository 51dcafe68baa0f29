#!/bin/bash
# one gpurun call: parity + timing of the attention op under several BG_ATTN_* settings (one process each: the knobs are
# read once per process)
#   gpurun --timeout 900 -- 'bash tools/attn_sweep.sh > gpurun_out/attn_sweep.log 2>&1'
run() { echo "=== $*"; env "$@" timeout 240 python tools/attn_check.py 2>&1 | tail -9; }
run B=64
run B=64 BG_ATTN_POLY=4
run B=64 BG_ATTN_POLY=5
run B=64 BG_ATTN_POLY=6
run B=256
run B=256 BG_ATTN_POLY=5
