"""clock64 trace of the attention softmax loop (debug build `make -C brepgen_b200/csrc trace`, loaded through BG_LIB):
per key block and query tile, cycles between the stamps placed in csrc/attn.cu (BG_TR)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("BG_LIB", os.path.join(ROOT, "brepgen_b200", "libbrepgen_trace.so"))
sys.path.insert(0, ROOT)
import torch
from brepgen_b200 import _ffi

B, L = 64, 4000
qkv = torch.randn(B * L, 2304, device="cuda", dtype=torch.float16)
out = torch.empty(B * L, 768, device="cuda", dtype=torch.float16)
lib = _ffi.lib()
for _ in range(3):
    _ffi.check(lib.bg_op_attention(qkv.data_ptr(), out.data_ptr(), B, L, None, 0, None, _ffi.current_stream()))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 64 * 8))()
raw = ctypes.CDLL(os.environ["BG_LIB"])
raw.bg_debug_attn_trace.argtypes = [ctypes.c_void_p]
assert raw.bg_debug_attn_trace(buf) == 0
tr = torch.tensor(list(buf)).view(2, 64, 8)
names = os.environ.get("BG_TR_NAMES", "0,6,1,2,7,3,4,5").split(",")
order = [int(x) for x in names]
print("env", {k: v for k, v in os.environ.items() if k.startswith("BG_ATTN")})
for t in range(2):
    print(f"tile {t}: per block: stamp deltas in order {order} (first column: block period)")
    for it in range(8, 20):
        row = tr[t, it]
        per = int(tr[t, it, 0] - tr[t, it - 1, 0])
        pts = [int(row[i]) for i in order if int(row[i]) != 0]
        d = [pts[i + 1] - pts[i] for i in range(len(pts) - 1)]
        print(f"  it={it:2d} period={per:5d}  deltas={d}  t0_rel_tile0={int(row[0] - tr[0, it, 0])}")
