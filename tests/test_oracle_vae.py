"""Structural invariants that pin the (diffusers-less) VAE decoder oracle -- SURVEY.md Appendix A.5 items 1, 3, 6."""
import math

import torch

from brepgen_b200.spec import CUBIC_UP_KERNEL, edge_decoder_spec, surf_decoder_spec
from brepgen_b200.synth import synth_state_dict
from oracle import vae as V


def test_parameter_counts_and_shapes():
    n = lambda sp: sum(math.prod(s) for k, s in sp if not k.endswith("up.kernel"))
    assert n(surf_decoder_spec()) == 49_485_583      # SD-VAE decoder 49 490 179 - 4 608 (3 vs 4 latent ch.) + 12
    assert n(edge_decoder_spec()) == 39_124_751
    sds, sde = synth_state_dict(surf_decoder_spec(), 3), synth_state_dict(edge_decoder_spec(), 3)
    with torch.no_grad():
        ys = V.surf_decode(sds, torch.randn(2, 3, 4, 4))
        ye = V.edge_decode(sde, torch.randn(3, 3, 4))
    assert ys.shape == (2, 3, 32, 32) and ye.shape == (3, 3, 32)      # consumed at sample.py:289-294
    assert torch.isfinite(ys).all() and torch.isfinite(ye).all()


def test_cubic_upsampler_partition_of_unity_and_ramp():
    k = torch.tensor(CUBIC_UP_KERNEL)
    assert abs(float(k[0::2].sum()) - 1) < 1e-6 and abs(float(k[1::2].sum()) - 1) < 1e-6
    x = torch.ones(1, 4, 8)
    assert torch.allclose(V.cubic_upsample1d(x, k), torch.ones(1, 4, 16), atol=1e-6)
    ramp = torch.arange(16.0).view(1, 1, 16)
    y = V.cubic_upsample1d(ramp, k)[0, 0, 8:24]          # interior: half-sample-phase ramp
    assert torch.allclose(y[1:] - y[:-1], torch.full((15,), 0.5), atol=1e-5)


def test_edge_decoder_oracle_matches_reference_wrapper_classes():
    """oracle edge_decode vs the outputs of the reference's OWN AutoencoderKL1DFastDecode / Decoder1D / UNetMidBlock1D /
    UpBlock1D (network.py:786-858, :188-299, :51-83, :30-48), instantiated with sample.py:86-97's arguments over module
    forms of the three diffusers leaves by tests/golden/make_golden_vae1d.py (which also checks the 193 state-dict keys and
    shapes of brepgen_b200.spec.edge_decoder_spec against that module tree).  Pins the wrapper the reference owns; the
    arithmetic inside the diffusers leaves stays unpinned."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_vae1d as G
    from brepgen_b200.spec import edge_decoder_spec
    from brepgen_b200.synth import synth_state_dict
    from oracle import vae as V
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae1d_golden.npz"))
    sd = synth_state_dict(edge_decoder_spec(), seed=2)
    for seed, n in ((0, 3), (1, 7)):
        with torch.no_grad():
            y = V.edge_decode(sd, G.inputs(seed, n)).numpy()
        ref = gold[f"s{seed}"]
        assert y.shape == ref.shape
        err = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
        assert err < 1e-5, err
    # the edge ENCODER wrapper the same way: AutoencoderKL1DFastEncode / Encoder1D (network.py:690-783, :86-185) with
    # trainer.py:841-852's arguments; 193 keys of edge_encoder_spec loaded strictly by the generator
    from brepgen_b200.spec import edge_encoder_spec
    sde = synth_state_dict(edge_encoder_spec(), seed=3)
    for seed, n in ((0, 2), (1, 5)):
        with torch.no_grad():
            y = V.edge_encode(sde, G.enc_inputs(seed, n)).numpy()
        ref = gold[f"enc_s{seed}"]
        assert y.shape == ref.shape
        err = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
        assert err < 1e-5, err
