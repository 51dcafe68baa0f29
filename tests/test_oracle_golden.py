"""Oracle (oracle/denoisers.py) vs golden vectors produced by the reference's own classes.

The fixtures in tests/golden/denoisers_golden.npz were written by tests/golden/make_golden.py, which
imports /root/reference/network.py (diffusers stubbed) and runs SurfPosNet/SurfZNet/EdgePosNet/EdgeZNet
(network.py:1066-1393) on synthetic weights.  Tolerance: fp32 re-association only (1e-5 of max|ref|).
"""
import os

import numpy as np
import pytest
import torch

from brepgen_b200.spec import denoiser_spec
from brepgen_b200.synth import synth_state_dict
from make_golden import case_inputs
from oracle import denoisers as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "denoisers_golden.npz"))


@pytest.mark.parametrize("kind", ["surfpos", "surfz", "edgepos", "edgez"])
@pytest.mark.parametrize("use_cf", [False, True])
def test_denoiser_matches_reference(kind, use_cf):
    sd = synth_state_dict(denoiser_spec(kind, use_cf), seed=7)
    for seed in (0, 1, 2):
        inp = case_inputs(kind, use_cf, seed)
        with torch.no_grad():
            y = O.FORWARDS[kind](sd, *inp.values()).numpy()
        ref = GOLD[f"{kind}|cf{int(use_cf)}|s{seed}"]
        assert y.shape == ref.shape
        err = np.abs(y - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (kind, use_cf, seed, err)


def test_sincos_matches_reference():
    y = O.sincos_embedding(torch.tensor([0, 1, 249, 999])).numpy()
    assert np.abs(y - GOLD["sincos|surfpos"]).max() < 1e-6
    # cos block first, then sin (network.py:1060)
    assert np.allclose(y[0, :384], 1.0) and np.allclose(y[0, 384:], 0.0)


def test_dedup_oracle_matches_reference_statements():
    """oracle/cascade.py dedup loops vs the outputs of the reference's own statements (sample.py:159-183, :242-261),
    executed verbatim by tests/golden/make_golden_dedup.py: packed boxes and both masks must be IDENTICAL."""
    from oracle.cascade import dedup_edges_np, dedup_surfaces_np
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "dedup_golden.npz"))
    n = len([k for k in gold.files if k.endswith("_surfPos_in")])
    assert n >= 5
    for i in range(n):
        sp = gold[f"c{i}_surfPos_in"]
        pos, mask = dedup_surfaces_np(sp, 0.08)
        assert np.array_equal(mask, gold[f"c{i}_surfMask"]), i
        assert np.array_equal(pos, gold[f"c{i}_surfPos_out"]), i
        em = dedup_edges_np(gold[f"c{i}_edgePos_in"], mask, 0.08)
        assert np.array_equal(em, gold[f"c{i}_edgeM"]), i
        assert mask.shape == sp.shape[:2] and em.shape == gold[f"c{i}_edgePos_in"].shape[:3]
