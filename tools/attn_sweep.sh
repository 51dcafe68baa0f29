#!/bin/bash
# one gpurun call: clock64 traces of the attention softmax loop
echo "=== PT=1 (default)"; timeout 240 python tools/attn_trace.py 2>&1 | tail -30
echo "=== PT=2"; BG_ATTN_PT=2 BG_TR_NAMES=0,6,1,2,3,4,5 timeout 240 python tools/attn_trace.py 2>&1 | tail -30
echo "=== PT=2 poly 0"; BG_ATTN_POLY=0 BG_ATTN_PT=2 BG_TR_NAMES=0,6,1,2,3,4,5 timeout 240 python tools/attn_trace.py 2>&1 | tail -30
