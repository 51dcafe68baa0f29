"""GPU parity of the drop-in VAE decoders against the CPU fp32 oracle (oracle/vae.py; parity unpinned at the diffusers
boundary, see its header).  Tolerance: 1e-3 relative L2 (fp16 tensor-core operands, fp32 accumulation / norms)."""
import pytest
import torch

from brepgen_b200.spec import edge_decoder_spec, surf_decoder_spec
from brepgen_b200.synth import synth_state_dict
from oracle import vae as V

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("N,chunk", [(1, 1024), (5, 2)])
def test_surface_decoder(N, chunk):
    from brepgen_b200.vae import AutoencoderKLFastDecode
    sd = synth_state_dict(surf_decoder_spec(), seed=5)
    m = AutoencoderKLFastDecode(in_channels=3, out_channels=3, block_out_channels=[128, 256, 512, 512], layers_per_block=2,
                                act_fn="silu", latent_channels=3, norm_num_groups=32, sample_size=512)
    missing = m.load_state_dict({**sd, "encoder.conv_in.bias": torch.zeros(128)}, strict=False)
    assert not missing.missing_keys and missing.unexpected_keys == ["encoder.conv_in.bias"]
    m = m.cuda().eval()
    m.chunk = chunk
    z = torch.randn(N, 3, 4, 4, generator=torch.Generator().manual_seed(N))
    with torch.no_grad():
        ref = V.surf_decode(sd, z)
        y = m(z.cuda()).cpu()
    assert y.shape == (N, 3, 32, 32) and torch.isfinite(y).all()
    err = rel_l2(y, ref)
    print(f"surface decoder N={N} rel_l2={err:.3e}")
    assert err < 1e-3, err


@pytest.mark.parametrize("N,chunk", [(3, 32768), (37, 16)])
def test_edge_decoder(N, chunk):
    from brepgen_b200.vae import AutoencoderKL1DFastDecode
    sd = synth_state_dict(edge_decoder_spec(), seed=6)
    m = AutoencoderKL1DFastDecode(in_channels=3, out_channels=3, block_out_channels=[128, 256, 512], layers_per_block=2,
                                  act_fn="silu", latent_channels=3, norm_num_groups=32, sample_size=512)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    m.chunk = chunk
    z = torch.randn(N, 3, 4, generator=torch.Generator().manual_seed(N))
    with torch.no_grad():
        ref = V.edge_decode(sd, z)
        y = m(z.cuda()).cpu()
    assert y.shape == (N, 3, 32) and torch.isfinite(y).all()
    err = rel_l2(y, ref)
    print(f"edge decoder N={N} rel_l2={err:.3e}")
    assert err < 1e-3, err


def test_cascade_with_decode_shapes():
    from brepgen_b200.models import NETS
    from brepgen_b200.sampler import Cascade, CascadeConfig
    from brepgen_b200.spec import denoiser_spec
    from brepgen_b200.vae import build_synthetic_decoders
    ms = {}
    for kind in NETS:
        m = NETS[kind](False)
        m.load_state_dict(synth_state_dict(denoiser_spec(kind, False), seed=11))
        ms[kind] = m.cuda().eval()
    sv, ev = build_synthetic_decoders("cuda")
    cfg = CascadeConfig(batch_size=2, num_surfaces=3, num_edges=4, schedule="ddpm", ddpm_steps=2, seed=1)
    out = Cascade(ms, sv, ev).run(cfg)
    assert out["surf_ncs"].shape == (2, 6, 32, 32, 3) and out["edge_ncs"].shape == (2, 6, 4, 32, 3)
    assert torch.isfinite(out["surf_ncs"]).all() and torch.isfinite(out["edge_ncs"]).all()
