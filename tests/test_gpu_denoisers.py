"""GPU parity of the four drop-in denoisers (brepgen_b200.models) through the C ABI.

 1. against the committed golden vectors produced by the reference's OWN classes (tests/golden/denoisers_golden.npz);
 2. against the CPU fp32 oracle (oracle/denoisers.py) on larger seeded shapes that exercise both attention variants
    (L <= 128 and L > 128), ragged key-padding masks, per-sample timesteps and classifier-free labels.
Bar (BASELINE.json north_star): <= 1e-3 relative (L2) vs the fp32 reference path, per forward.
"""
import os

import numpy as np
import pytest
import torch

from brepgen_b200.spec import denoiser_spec
from brepgen_b200.synth import synth_state_dict
from make_golden import case_inputs
from oracle import denoisers as O

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "denoisers_golden.npz"))
TOL = 1e-3


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _model(kind, use_cf, seed=7):
    from brepgen_b200.models import NETS
    m = NETS[kind](use_cf)
    sd = synth_state_dict(denoiser_spec(kind, use_cf), seed=seed)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def _cuda(v):
    return v.cuda() if torch.is_tensor(v) else v


@pytest.mark.parametrize("kind", ["surfpos", "surfz", "edgepos", "edgez"])
@pytest.mark.parametrize("use_cf", [False, True])
def test_golden(kind, use_cf):
    m, _ = _model(kind, use_cf)
    for seed in (0, 1, 2):
        inp = case_inputs(kind, use_cf, seed)
        with torch.no_grad():
            y = m(*[_cuda(v) for v in inp.values()]).cpu()
        ref = torch.from_numpy(GOLD[f"{kind}|cf{int(use_cf)}|s{seed}"])
        assert y.shape == ref.shape
        # padded rows are discarded downstream (sample.py:245,284): with token compaction (the default) they are not computed
        # at all (0); they must be finite, and the comparison is over the valid tokens
        assert torch.isfinite(y).all()
        mask = inp.get("surf_mask", inp.get("mask"))
        if mask is not None:
            keep = ~mask
            if kind == "edgepos":
                keep = keep[..., None].expand(y.shape[:3])
            y, ref = y[keep], ref[keep]
        err = rel_l2(y, ref)
        print(f"golden {kind} cf={use_cf} seed={seed} rel_l2={err:.3e}")
        assert err < TOL, err


def _big_inputs(kind, use_cf, B, S, E, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    nvalid = torch.randint(max(1, S // 3), S + 1, (B,), generator=g)
    surf_mask = torch.arange(S)[None, :] >= nvalid[:, None]
    edge_mask = surf_mask[..., None].repeat(1, 1, max(E, 1)) | (torch.rand(B, S, max(E, 1), generator=g) < 0.3)
    edge_mask[:, :, 0] = surf_mask
    label = torch.randint(0, 11, (B, 1), generator=g) if use_cf else None
    t = torch.randint(0, 1000, (B,), generator=g) if seed % 2 else torch.tensor([37 * seed % 1000])
    if kind == "surfpos":
        return dict(surfPos=r(B, S, 6), timesteps=t, class_label=label)
    if kind == "surfz":
        return dict(surfZ=r(B, S, 48), timesteps=t, surfPos=r(B, S, 6), surf_mask=surf_mask, class_label=label)
    if kind == "edgepos":
        return dict(edgePos=r(B, S, E, 6), timesteps=t, surfPos=r(B, S, 6), surfZ=r(B, S, 48), mask=surf_mask, class_label=label)
    return dict(edge=r(B, S, E, 18), timesteps=t, edgePos=r(B, S, E, 6), surfPos=r(B, S, 6), surfZ=r(B, S, 48),
                mask=edge_mask, class_label=label)


BIG = [("surfpos", False, 4, 50, 0, 2), ("surfpos", True, 3, 100, 0, 3), ("surfz", False, 3, 100, 0, 4),
       ("surfz", True, 2, 30, 0, 5), ("edgepos", False, 2, 12, 20, 6), ("edgepos", True, 1, 30, 30, 7),
       ("edgez", False, 2, 10, 30, 8), ("edgez", True, 1, 25, 40, 9)]


@pytest.mark.parametrize("kind,use_cf,B,S,E,seed", BIG)
def test_vs_oracle(kind, use_cf, B, S, E, seed):
    m, sd = _model(kind, use_cf)
    inp = _big_inputs(kind, use_cf, B, S, E, seed)
    with torch.no_grad():
        ref = O.FORWARDS[kind](sd, *inp.values())
        y = m(*[_cuda(v) for v in inp.values()]).cpu()
    assert torch.isfinite(y).all()
    # compare on valid tokens only (outputs of padded tokens are discarded by the cascade)
    mask = inp.get("surf_mask", inp.get("mask"))
    if mask is not None:
        keep = ~mask
        if kind == "edgepos":
            keep = keep[..., None].expand(B, S, E)
        y, ref = y[keep], ref[keep]
    err = rel_l2(y, ref)
    print(f"oracle {kind} cf={use_cf} B={B} S={S} E={E} rel_l2={err:.3e}")
    assert err < TOL, err


@pytest.mark.parametrize("kind,use_cf,B,S,E,seed", [BIG[0], BIG[6]])
def test_precision_modes(kind, use_cf, B, S, E, seed):
    """precision 0 (plain fp16 operands) ~ the reference's own fp16-autocast error; 1 and 2 tighten it."""
    m, sd = _model(kind, use_cf)
    inp = _big_inputs(kind, use_cf, B, S, E, seed)
    with torch.no_grad():
        ref = O.FORWARDS[kind](sd, *inp.values())
    mask = inp.get("surf_mask", inp.get("mask"))
    errs = []
    for prec in (0, 1, 2):
        m.precision = prec
        with torch.no_grad():
            y = m(*[_cuda(v) for v in inp.values()]).cpu()
        a, b = (y[~mask], ref[~mask]) if mask is not None else (y, ref)
        errs.append(rel_l2(a, b))
    print(f"precision modes {kind}: " + " ".join(f"p{i}={e:.3e}" for i, e in enumerate(errs)))
    assert errs[0] < 2.5e-3 and errs[1] < TOL and errs[2] < TOL
    assert errs[2] <= errs[0]


def test_forward_is_deterministic_and_repack_on_load():
    m, sd = _model("surfz", False)
    inp = {k: _cuda(v) for k, v in _big_inputs("surfz", False, 2, 40, 0, 4).items()}
    y1 = m(*inp.values())
    y2 = m(*inp.values())
    assert torch.equal(y1, y2)
    sd2 = synth_state_dict(denoiser_spec("surfz", False), seed=8)
    m.load_state_dict(sd2)
    y3 = m(*inp.values())
    assert not torch.allclose(y1, y3)


def test_cpu_input_raises():
    from brepgen_b200.models import SurfPosNet
    with pytest.raises(RuntimeError):
        SurfPosNet(False)(torch.zeros(1, 4, 6), torch.tensor([3]), None)
