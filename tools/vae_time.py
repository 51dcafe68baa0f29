"""decode time of the two VAE decoders at the cascade's shapes: B B-reps = B * 100 faces (surface) + B * 4000 edges"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brepgen_b200.vae import build_synthetic_decoders

B = int(os.environ.get("B", 64))
dev = torch.device("cuda:0")
surf, edge = build_synthetic_decoders(dev)
zs = torch.randn(B * 100, 3, 4, 4, device=dev)
ze = torch.randn(B * 4000, 3, 4, device=dev)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ms_s = timed(lambda: surf(zs))
    ms_e = timed(lambda: edge(ze))
gf = (B * 100 * 9.69 + B * 4000 * 0.416)
print(f"B={B}: surface {ms_s:.1f} ms, edge {ms_e:.1f} ms, total {ms_s + ms_e:.1f} ms = {gf / (ms_s + ms_e):.0f} TF/s algorithmic "
      f"({(ms_s + ms_e) * 256 / B / 1000:.2f} s per 256 B-reps)")
