#!/bin/bash
set -x
BREPGEN_B200_VAE_TERMS=2 timeout 600 python -m pytest tests/test_gpu_vae.py -q -s 2>&1 | grep -E "rel_l2|passed|failed|Error" | head -20
timeout 600 python -m pytest tests/test_gpu_vae.py -q -s 2>&1 | grep -E "rel_l2|passed|failed" | head -20
cat > /tmp/vt.py <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from brepgen_b200.vae import build_synthetic_decoders
sv, ev = build_synthetic_decoders(torch.device('cuda'))
zs = torch.randn(64 * 100, 3, 4, 4, device='cuda'); ze = torch.randn(64 * 4000, 3, 4, device='cuda')
for _ in range(2): sv(zs); ev(ze)
torch.cuda.synchronize(); t0 = time.time(); sv(zs); torch.cuda.synchronize(); t1 = time.time(); ev(ze); torch.cuda.synchronize(); t2 = time.time()
print(f"decode B=64: surface {1e3*(t1-t0):.1f} ms, edge {1e3*(t2-t1):.1f} ms")
PY
BREPGEN_B200_VAE_TERMS=2 python /tmp/vt.py
BREPGEN_B200_VAE_TERMS=3 python /tmp/vt.py
