"""CPU tests: the 2-D VAE leaves of oracle/vae.py against the known answers of diffusers' OWN layer tests.

diffusers (requirements.txt:5 of the reference pins 0.27) is absent from the image; its tests/models/test_layers_utils.py
publishes expected output slices for layers built under torch.manual_seed(0) with torch's default initialisers.  The torch
modules below are created in the order diffusers' constructors create theirs, so they draw the same parameters from the
same generator; the arithmetic under test is the oracle's.  Expected slices (output[0, -1, -3:, -3:]) and the 1e-3 tolerance
are diffusers'."""
import torch
import torch.nn as nn

from oracle import vae as V


def _gn(prefix, C):
    return {prefix + ".weight": torch.ones(C), prefix + ".bias": torch.zeros(C)}


def _conv(prefix, m):
    return {prefix + ".weight": m.weight.data, prefix + ".bias": m.bias.data}


def _check(out, expected):
    got = out[0, -1, -3:, -3:].flatten()
    assert torch.allclose(got, torch.tensor(expected), atol=1e-3), got


def _resnet_case(shortcut):
    torch.manual_seed(0)
    sample, temb = torch.randn(1, 32, 64, 64), torch.randn(1, 128)
    # ResnetBlock2D.__init__(in_channels=32, temb_channels=128): norm1, conv1, time_emb_proj, norm2, conv2[, conv_shortcut]
    conv1, tproj, conv2 = nn.Conv2d(32, 32, 3, padding=1), nn.Linear(128, 32), nn.Conv2d(32, 32, 3, padding=1)
    sd = {**_gn("r.norm1", 32), **_gn("r.norm2", 32), **_conv("r.conv1", conv1), **_conv("r.conv2", conv2)}
    if shortcut:
        sd.update(_conv("r.conv_shortcut", nn.Conv2d(32, 32, 1)))
    with torch.no_grad():
        return V._resnet2d(sd, "r", sample, temb_add=tproj(torch.nn.functional.silu(temb))[:, :, None, None])


def test_resnet_block_default():          # ResnetBlock2DTests.test_resnet_default
    _check(_resnet_case(False), [-1.9010, -0.2974, -0.8245, -1.3533, 0.8742, -0.9645, -2.0584, 1.3387, -0.4746])


def test_resnet_block_conv_shortcut():    # ResnetBlock2DTests.test_restnet_with_use_in_shortcut
    _check(_resnet_case(True), [0.2226, -1.0791, -0.1629, 0.3659, -0.2889, -1.2376, 0.0582, 0.9206, 0.0044])


def test_attention_block_default():       # AttentionBlockTests.test_attention_block_default (32 heads of dimension 1)
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 64, 64)
    q, k, v, o = (nn.Linear(32, 32) for _ in range(4))     # group_norm, query, key, value, proj_attn
    sd = {**_gn("a.group_norm", 32), **_conv("a.to_q", q), **_conv("a.to_k", k), **_conv("a.to_v", v), **_conv("a.to_out.0", o)}
    with torch.no_grad():
        out = V._attn2d(sd, "a", sample, n_head=32)
    _check(out, [-1.4975, -0.0038, -0.7847, -1.4567, 1.1220, -0.8962, -1.7394, 1.1319, -0.5427])


def test_upsample_with_conv():            # Upsample2DBlockTests.test_upsample_with_conv
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 32, 32)
    sd = _conv("u.conv", nn.Conv2d(32, 32, 3, padding=1))
    with torch.no_grad():
        out = V._upsample2d(sd, "u", sample)
    assert out.shape == (1, 32, 64, 64)
    _check(out, [0.7145, 1.3773, 0.3492, 0.8448, 1.0839, -0.3341, 0.5956, 0.1250, -0.4841])


# ---- block level (diffusers tests/models/unets/test_unet_2d_blocks.py over test_unet_blocks_common.py: torch.manual_seed(0),
# hidden_states randn(4, 32, 32, 32) [, temb randn(4, 128)], THEN the block is constructed from the same generator)
def test_up_decoder_block():              # UpDecoderBlock2DTests.test_output: ResnetBlock2D (no time embedding) -> Upsample2D(conv)
    torch.manual_seed(0)
    x = torch.randn(4, 32, 32, 32)
    c1, c2, up = nn.Conv2d(32, 32, 3, padding=1), nn.Conv2d(32, 32, 3, padding=1), nn.Conv2d(32, 32, 3, padding=1)
    sd = {**_gn("r.norm1", 32), **_gn("r.norm2", 32), **_conv("r.conv1", c1), **_conv("r.conv2", c2), **_conv("u.conv", up)}
    with torch.no_grad():
        out = V._upsample2d(sd, "u", V._resnet2d(sd, "r", x))       # exactly the VAE decoder's up block (temb is None there)
    assert out.shape == (4, 32, 64, 64)
    _check(out, [0.4404, 0.1998, -0.9886, -0.3320, -0.3128, -0.7034, -0.6955, -0.2338, -0.3137])


def test_unet_mid_block():                # UNetMidBlock2DTests.test_output: resnet -> attention (heads of dimension 1) -> resnet
    torch.manual_seed(0)
    x, temb = torch.randn(4, 32, 32, 32), torch.randn(4, 128)
    mk = lambda: (nn.Conv2d(32, 32, 3, padding=1), nn.Linear(128, 32), nn.Conv2d(32, 32, 3, padding=1))
    r0 = mk()                             # UNetMidBlock2D.__init__: resnets[0], attentions[0], resnets[1]
    q, k, v, o = (nn.Linear(32, 32) for _ in range(4))
    r1 = mk()
    sd = {**_gn("a.group_norm", 32), **_conv("a.to_q", q), **_conv("a.to_k", k), **_conv("a.to_v", v), **_conv("a.to_out.0", o)}
    for name, (c1, _, c2) in (("r0", r0), ("r1", r1)):
        sd.update({**_gn(name + ".norm1", 32), **_gn(name + ".norm2", 32), **_conv(name + ".conv1", c1), **_conv(name + ".conv2", c2)})
    silu = torch.nn.functional.silu
    with torch.no_grad():
        y = V._resnet2d(sd, "r0", x, temb_add=r0[1](silu(temb))[:, :, None, None])
        y = V._attn2d(sd, "a", y, n_head=32)
        y = V._resnet2d(sd, "r1", y, temb_add=r1[1](silu(temb))[:, :, None, None])
    _check(y, [-0.1062, 1.7248, 0.3494, 1.4569, -0.0910, -1.2421, -0.9984, 0.6736, 1.0028])


def test_downsample_with_conv():          # Downsample2DBlockTests.test_downsample_with_conv (default padding = 1)
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 64, 64)
    sd = _conv("d.conv", nn.Conv2d(32, 32, 3, stride=2, padding=1))
    with torch.no_grad():
        out = V._downsample2d(sd, "d", sample, padding=1)
    assert out.shape == (1, 32, 32, 32)
    _check(out, [0.9267, 0.5878, 0.3337, 1.2321, -0.1191, -0.3984, -0.7532, -0.0715, -0.3913])


def test_down_encoder_block():            # DownEncoderBlock2DTests.test_output: ResnetBlock2D (no time embedding) -> Downsample2D
    torch.manual_seed(0)
    x = torch.randn(4, 32, 32, 32)
    c1, c2, d = nn.Conv2d(32, 32, 3, padding=1), nn.Conv2d(32, 32, 3, padding=1), nn.Conv2d(32, 32, 3, stride=2, padding=1)
    sd = {**_gn("r.norm1", 32), **_gn("r.norm2", 32), **_conv("r.conv1", c1), **_conv("r.conv2", c2), **_conv("d.conv", d)}
    with torch.no_grad():
        out = V._downsample2d(sd, "d", V._resnet2d(sd, "r", x), padding=1)
    assert out.shape == (4, 32, 16, 16)
    _check(out, [1.1102, 0.5302, 0.4872, -0.0023, -0.8042, 0.0483, -0.3489, -0.5632, 0.7626])
