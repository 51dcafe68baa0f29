"""one-process parity + timing check of the attention op under the current BG_ATTN_* environment (used for variants)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brepgen_b200 import _ffi

torch.backends.cuda.matmul.allow_tf32 = False


def ref(qkv, B, L, mask):
    q, k, v = qkv.float().view(B, L, 3, 12, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s.masked_fill(mask.view(B, 1, 1, L), float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, 768)


def run(qkv, B, L, mask):
    out = torch.full((B * L, 768), float("nan"), device="cuda", dtype=torch.float16)
    nkb = (L + 127) // 128
    scr = torch.zeros(B * (5 * nkb + 1), dtype=torch.int32, device="cuda")
    _ffi.check(_ffi.lib().bg_op_attention(qkv.data_ptr(), out.data_ptr(), B, L, _ffi.ptr(mask), 1 if mask is not None else 0,
                                        scr.data_ptr(), _ffi.current_stream()), "attention")
    torch.cuda.synchronize()
    return out


env = {k: v for k, v in os.environ.items() if k.startswith("BG_ATTN")}
g = torch.Generator(device="cuda").manual_seed(7)
for name, B, L in (("plain", 1, 4000), ("ragged", 8, 1000), ("growing", 3, 1500), ("shrinking", 3, 1500)):
    qkv = torch.randn(B * L, 2304, generator=g, device="cuda") * 1.5
    mask = None
    if name == "ragged":
        nv = torch.randint(1, L + 1, (B,), generator=g, device="cuda")
        nv[0] = L
        mask = (torch.arange(L, device="cuda")[None] >= nv[:, None]) | (torch.rand(B, L, generator=g, device="cuda") < 0.1)
        mask[:, 0] = False
    if name in ("growing", "shrinking"):      # key norms change by 6x along the sequence: exercises the rescale / redo paths
        ramp = torch.linspace(1, 6, L, device="cuda") if name == "growing" else torch.linspace(6, 1, L, device="cuda")
        qkv.view(B, L, 2304)[:, :, 768:1536] *= ramp[None, :, None]
    qkv = qkv.half()
    o, r = run(qkv, B, L, mask).float(), ref(qkv, B, L, mask)
    err = float((o.double() - r.double()).norm() / r.double().norm())
    print(f"{env} {name} B={B} L={L}: rel_l2={err:.3e} finite={bool(torch.isfinite(o).all())}", flush=True)

B, L = int(os.environ.get("B", 64)), 4000
qkv = torch.randn(B * L, 2304, device="cuda", dtype=torch.float16)
mask = torch.zeros(B, L, dtype=torch.bool, device="cuda")
out = torch.empty(B * L, 768, device="cuda", dtype=torch.float16)
scr = torch.zeros(B * 170, dtype=torch.int32, device="cuda")
go = lambda: _ffi.check(_ffi.lib().bg_op_attention(qkv.data_ptr(), out.data_ptr(), B, L, mask.data_ptr(), 1, scr.data_ptr(), _ffi.current_stream()))
for _ in range(3): go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): go()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{env} timing B={B} L={L}: {ms:.3f} ms  {B*3072*L*L/ms/1e9:.0f} TF/s", flush=True)
