import os, sys
sys.path.insert(0, os.getcwd())
import torch
from brepgen_b200.vae import build_synthetic_decoders
dev = torch.device("cuda:0")
surf, edge = build_synthetic_decoders(dev)
B = 16
zs = torch.randn(B * 100, 3, 4, 4, device=dev)
ze = torch.randn(B * 4000, 3, 4, device=dev)
with torch.no_grad():
    which = os.environ.get("WHICH", "edge")
    (edge(ze) if which == "edge" else surf(zs))
torch.cuda.synchronize()
