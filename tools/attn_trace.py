"""debug: run one attention launch (B=8, L=4000) with the BG_ATTN_TRACE build and print per-phase clock64 deltas"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brepgen_b200 import _ffi
B, L = 8, 4000
qkv = torch.randn(B * L, 2304, device="cuda", dtype=torch.float16)
out = torch.empty(B * L, 768, device="cuda", dtype=torch.float16)
for _ in range(3):
    _ffi.check(_ffi.lib().bg_op_attention(qkv.data_ptr(), out.data_ptr(), B, L, None, 0, None, _ffi.current_stream()))
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 32 * 8))()
lib = ctypes.CDLL(_ffi.LIB_PATH)
print("read", lib.bg_debug_read_trace(buf))
import numpy as np
a = np.array(buf).reshape(2, 32, 8)
t0 = a[:, 0, 0].min()
names = ["wait_s", "ldtm", "max", "wait_pv", "barsync", "exp", "tail"]
for it in range(8, 20):
    for t in range(2):
        d = np.diff(a[t, it])
        print(f"t={t} it={it:2d} start={a[t, it, 0] - t0:7d} period={a[t, it, 0] - a[t, it - 1, 0]:5d} " +
              " ".join(f"{n}={v}" for n, v in zip(names, d)))
