"""GPU: mask-aware token compaction (BgDenoiserArgs.compact) against the dense-layout path of
the same kernels: identical semantics on valid tokens (padded keys are never attended to, network.py:1268,1387-1390; padded
outputs are discarded downstream, sample.py:245,284), outputs of padded tokens exactly 0.  The reference-golden tests
(test_gpu_l4000.py, test_gpu_denoisers.py) run with compaction on (the default), so this file pins compact == dense."""
import pytest
import torch

from brepgen_b200.spec import denoiser_spec
from brepgen_b200.synth import synth_state_dict

pytestmark = pytest.mark.gpu


def _model(kind, use_cf):
    from brepgen_b200.models import NETS
    m = NETS[kind](use_cf)
    m.load_state_dict(synth_state_dict(denoiser_spec(kind, use_cf), seed=7))
    return m.cuda().eval()


CASES = [("surfz", False, 3, 100, 0), ("surfz", True, 5, 30, 0), ("edgepos", False, 2, 30, 30), ("edgepos", True, 3, 100, 40),
         ("edgez", False, 3, 100, 40), ("edgez", True, 2, 17, 9)]


@pytest.mark.parametrize("kind,use_cf,B,S,E", CASES)
def test_compact_equals_dense_on_valid_tokens(kind, use_cf, B, S, E):
    g = torch.Generator().manual_seed(B * 1000 + S * 10 + E)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    nvalid = torch.randint(1, S + 1, (B,), generator=g)
    nvalid[0] = S                                               # one sample without padded faces
    if B > 1:
        nvalid[1] = 1                                           # and one with a single valid face
    surf_mask = (torch.arange(S)[None, :] >= nvalid[:, None]).cuda()
    label = torch.randint(0, 11, (B, 1), generator=g).cuda() if use_cf else None
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    if kind == "surfz":
        args, valid = (r(B, S, 48), t, r(B, S, 6), surf_mask, label), ~surf_mask
    elif kind == "edgepos":
        args, valid = (r(B, S, E, 6), t, r(B, S, 6), r(B, S, 48), surf_mask, label), ~surf_mask[..., None].expand(B, S, E)
    else:
        em = surf_mask[..., None].repeat(1, 1, E) | (torch.rand(B, S, E, generator=g) < 0.4).cuda()
        em[:, 0, 0] = False
        args, valid = (r(B, S, E, 18), t, r(B, S, E, 6), r(B, S, 6), r(B, S, 48), em, label), ~em
    m = _model(kind, use_cf)
    with torch.no_grad():
        m.compact = 0
        dense = m(*args)
        m.compact = 1
        comp = m(*args)
        comp2 = m(*args)
    torch.cuda.synchronize()
    assert torch.equal(comp, comp2)                             # deterministic
    assert torch.isfinite(comp).all()
    assert float(comp[~valid].abs().max()) == 0.0 if (~valid).any() else True
    err = float((comp[valid].double() - dense[valid].double()).norm() / dense[valid].double().norm())
    print(f"compaction {kind} cf={use_cf} B={B} S={S} E={E}: valid {int(valid.sum())}/{valid.numel()} tokens, rel_l2 vs dense {err:.2e}")
    assert err < 3e-4, err


def test_compaction_first_use_of_an_uninitialised_workspace():
    """the rows behind the last valid token are read by the last key tile (masked keys): they must not leak NaN"""
    m = _model("edgez", False)
    B, S, E = 2, 20, 13
    g = torch.Generator().manual_seed(1)
    em = (torch.rand(B, S, E, generator=g) < 0.5).cuda()
    em[:, 0, 0] = False
    args = (torch.randn(B, S, E, 18, generator=g).cuda(), torch.tensor([7]).cuda(), torch.randn(B, S, E, 6, generator=g).cuda(),
            torch.randn(B, S, 6, generator=g).cuda(), torch.randn(B, S, 48, generator=g).cuda(), em, None)
    with torch.no_grad():
        m(*args)                                               # allocates the workspace
        next(iter(m._ws.values())).view(torch.float16).fill_(float("nan"))
        y = m(*args)
    assert torch.isfinite(y).all()
