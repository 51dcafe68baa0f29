"""time the edge-stage attention kernel alone (CUDA events), B x L = 64 x 4000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brepgen_b200 import _ffi
B, L = int(os.environ.get("B", 64)), 4000
qkv = torch.randn(B * L, 2304, device="cuda", dtype=torch.float16)
out = torch.empty(B * L, 768, device="cuda", dtype=torch.float16)
mask = torch.zeros(B, L, dtype=torch.bool, device="cuda")
scr = torch.zeros(B * 170, dtype=torch.int32, device="cuda")
run = lambda: _ffi.check(_ffi.lib().bg_op_attention(qkv.data_ptr(), out.data_ptr(), B, L, mask.data_ptr(), 1, scr.data_ptr(), _ffi.current_stream()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{os.path.basename(_ffi.LIB_PATH)} PS={os.environ.get('BG_ATTN_PS', '0')} attention B={B} L={L}: {ms:.3f} ms  {B*3072*L*L/ms/1e9:.0f} TF/s")
