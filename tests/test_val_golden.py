"""Forward-only validation of the four LDM trainers (brepgen_b200/validation.py = the reference's test_val() loop bodies,
trainer.py:395-403, 579-596, 774-791, 997-1019) against tests/golden/val_golden.npz, the loss sums produced by exec()'ing
those reference lines verbatim (tests/golden/make_golden_val.py).  CPU: the same oracle callables stand in for the models /
encoders / scheduler, so only the restated glue is under test (tight tolerance).  GPU: the product models, encoders and
scheduler carry the work through the C ABI (1e-3)."""
import os

import numpy as np
import pytest
import torch

from brepgen_b200 import validation as VAL
from make_golden_val import CF_LABEL, SEEDS, Z_SCALED, oracle_callables, state_dicts, state_dicts_cf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "val_golden.npz")


def _load(device="cpu"):
    g = np.load(GOLD)
    t = {k: torch.from_numpy(g[k]).to(device) for k in ("surfPos", "surfPnt", "edgePos", "edgePnt", "vertPos", "surf_mask", "edge_mask")}
    return g, t


def _run_all(models, surf_vae, edge_vae, sched, t, label=None):
    out = {}
    lab = (lambda: None) if label is None else (lambda: label.clone())   # the reference may modify labels in place (is_train)
    torch.manual_seed(SEEDS["surfpos"])
    out["surfpos"] = VAL.surfpos_val_losses(models["surfpos"], sched, t["surfPos"], lab(), rng_device="cpu")
    torch.manual_seed(SEEDS["surfz"])
    out["surfz"] = VAL.surfz_val_losses(models["surfz"], surf_vae, sched, t["surfPos"], t["surfPnt"], t["surf_mask"], lab(), Z_SCALED,
                                        rng_device="cpu")
    torch.manual_seed(SEEDS["edgepos"])
    out["edgepos"] = VAL.edgepos_val_losses(models["edgepos"], surf_vae, sched, t["edgePos"], t["surfPnt"], t["surfPos"],
                                            t["surf_mask"], lab(), Z_SCALED, rng_device="cpu")
    torch.manual_seed(SEEDS["edgez"])
    out["edgez"] = VAL.edgez_val_losses(models["edgez"], surf_vae, edge_vae, sched, t["edgePnt"], t["edgePos"], t["edge_mask"],
                                        t["surfPnt"], t["surfPos"], t["vertPos"], lab(), Z_SCALED, rng_device="cpu")
    return out


@pytest.mark.parametrize("cf", [False, True])
def test_validation_glue_matches_reference_statements(cf):
    from oracle.schedulers import DDPMOracle
    g, t = _load()
    models, surf_vae, edge_vae = oracle_callables(state_dicts_cf() if cf else state_dicts())
    out = _run_all(models, surf_vae, edge_vae, DDPMOracle(), t, CF_LABEL if cf else None)
    for name, got in out.items():
        ref = g[f"loss_{name}_cf" if cf else f"loss_{name}"]
        assert len(got) == len(ref)
        assert np.allclose(np.asarray(got), ref, rtol=1e-5, atol=0), (name, got, ref)


@pytest.mark.gpu
def test_validation_on_the_product_path_matches_reference_statements():
    from brepgen_b200.models import EdgePosNet, EdgeZNet, SurfPosNet, SurfZNet
    from brepgen_b200.schedulers import DDPMScheduler
    from brepgen_b200.vae import AutoencoderKL1DFastEncode, AutoencoderKLFastEncode
    g, t = _load("cuda")
    sd = state_dicts()
    models = {}
    for name, cls in (("surfpos", SurfPosNet), ("surfz", SurfZNet), ("edgepos", EdgePosNet), ("edgez", EdgeZNet)):
        m = cls(False)
        m.load_state_dict(sd[name])
        models[name] = m.cuda().eval()
    es = AutoencoderKLFastEncode(block_out_channels=[128, 256, 512, 512])
    es.load_state_dict({**sd["surf_enc"], "decoder.conv_in.bias": torch.zeros(512)}, strict=False)
    ee = AutoencoderKL1DFastEncode(block_out_channels=[128, 256, 512])
    ee.load_state_dict(sd["edge_enc"], strict=False)
    sched = DDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon", beta_start=0.0001,
                          beta_end=0.02, clip_sample=False)       # trainer.py:285-292
    out = _run_all(models, es.cuda().eval(), ee.cuda().eval(), sched, t)
    for name, got in out.items():
        ref = g[f"loss_{name}"]
        err = np.abs(np.asarray(got) - ref) / np.abs(ref)
        print(f"validation {name}: max rel diff of the loss sums {err.max():.2e}")
        assert err.max() < 1e-3, (name, got, ref)
