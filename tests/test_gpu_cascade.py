"""GPU parity of the stage glue and the whole (short) cascade against the CPU oracle.

  * de-duplication kernels: bit-exact vs the numpy loops of sample.py:159-183 / :242-261 (oracle/cascade.py), including
    corner-swapped matches, values on the 4-decimal rounding boundary, fully padded faces, S and E up to the ABC sizes;
  * scheduler kernels: DDPM chain with injected noise and the full 209-step PNDM chain (PRK + PLMS) vs the oracle;
  * a short cascade (4 DDPM steps per stage and the shipped hybrid truncated by the oracle's own tables) vs
    oracle/cascade.py with identical initial and step noise: masks identical, tensors within 1e-3 relative.
"""
import numpy as np
import pytest
import torch

from brepgen_b200.spec import denoiser_spec
from brepgen_b200.synth import synth_state_dict
from oracle import cascade as OC
from oracle.schedulers import DDPMOracle, PNDMOracle

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _boxes(g, *shape):
    """random boxes with many near-duplicates, some corner-swapped, some exactly on rounding boundaries"""
    base = torch.randn(*shape[:-1], 6, generator=g)
    n = shape[-2]
    flat = base.view(-1, n, 6)
    for i in range(1, n):
        # ~45 % of the slots re-use an earlier slot, perturbed by up to +-0.1 per coordinate (threshold is 0.08, so
        # both outcomes occur), 30 % of those with the two corners swapped
        src = torch.randint(0, i, (flat.shape[0],), generator=g)
        cp = flat[torch.arange(flat.shape[0]), src] + (torch.rand(flat.shape[0], 6, generator=g) - 0.5) * 0.2 \
            * torch.rand(flat.shape[0], 1, generator=g)
        sw = torch.rand(flat.shape[0], generator=g) < 0.3
        cp[sw] = torch.cat([cp[sw][:, 3:], cp[sw][:, :3]], -1)
        use = torch.rand(flat.shape[0], generator=g) < 0.45
        flat[use, i] = cp[use]
    edge = torch.rand(shape[:-1], generator=g) < 0.1
    base[edge] = torch.round(base[edge] * 1e4) / 1e4 + 5e-5      # values on the np.round(.,4) tie boundary
    return base.float()


@pytest.mark.parametrize("B,S", [(1, 1), (3, 7), (4, 50), (2, 100)])
def test_dedup_surfaces_bit_exact(B, S):
    from brepgen_b200.sampler import dedup_surfaces
    g = torch.Generator().manual_seed(B * 100 + S)
    pos = _boxes(g, B, S, 6)
    ref_pos, ref_mask = OC.dedup_surfaces_np(pos.numpy(), np.float32(0.08))
    out, mask = dedup_surfaces(pos.cuda(), 0.08)
    assert np.array_equal(mask.cpu().numpy(), ref_mask)
    assert np.array_equal(out.cpu().numpy(), ref_pos)
    assert ref_mask.sum() > 0 or S == 1      # the generator does produce duplicates


@pytest.mark.parametrize("B,S,E", [(1, 2, 3), (2, 9, 30), (2, 100, 40)])
def test_dedup_edges_bit_exact(B, S, E):
    from brepgen_b200.sampler import dedup_edges
    g = torch.Generator().manual_seed(B * 1000 + S * 10 + E)
    pos = _boxes(g, B, S, E, 6)
    nvalid = torch.randint(1, S + 1, (B,), generator=g)
    smask = torch.arange(S)[None] >= nvalid[:, None]
    ref = OC.dedup_edges_np(pos.numpy(), smask.numpy(), np.float32(0.08))
    m = dedup_edges(pos.cuda(), smask.cuda(), 0.08)
    assert np.array_equal(m.cpu().numpy(), ref)


def test_ddpm_chain_matches_oracle():
    from brepgen_b200.schedulers import DDPMScheduler
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 37, 6, generator=g)
    sched, orc = DDPMScheduler(clip_sample=True, clip_sample_range=3), DDPMOracle()
    for n, sl in ((1000, slice(-250, None, 25)), (8, slice(None))):
        sched.set_timesteps(n), orc.set_timesteps(n)
        xo, xg = x.clone(), x.clone().cuda()
        for t in sched.timesteps[sl]:
            eps = torch.tanh(xo * 0.7) + 0.1
            nz = torch.randn(x.shape, generator=g)
            xo = orc.step(eps, int(t), xo, nz)
            xg = sched.step((torch.tanh(xg * 0.7) + 0.1), t, xg, noise=nz).prev_sample
        assert rel_l2(xg.cpu(), xo) < 1e-5


def test_pndm_chain_matches_oracle():
    from brepgen_b200.schedulers import PNDMScheduler
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 11, 48, generator=g)
    sched, orc = PNDMScheduler(), PNDMOracle()
    sched.set_timesteps(200), orc.set_timesteps(200)
    xo, xg = x.clone(), x.clone().cuda()
    for i, t in enumerate(sched.timesteps):
        xo = orc.step(torch.tanh(xo * 0.7) + 0.1, int(t), xo)
        xg = sched.step(torch.tanh(xg * 0.7) + 0.1, t, xg).prev_sample
        if i in (0, 11, 12, 50, 157, 208):
            assert rel_l2(xg.cpu(), xo) < 2e-5, i
    assert sched.counter == 209


def _models(use_cf):
    from brepgen_b200.models import NETS
    ms, sds = {}, {}
    for kind in NETS:
        sds[kind] = synth_state_dict(denoiser_spec(kind, use_cf), seed=11)
        m = NETS[kind](use_cf)
        m.load_state_dict(sds[kind])
        ms[kind] = m.cuda().eval()
    return ms, sds


@pytest.mark.parametrize("use_cf,schedule", [(False, "ddpm"), (True, "ddpm")])
def test_short_cascade_matches_oracle(use_cf, schedule):
    from brepgen_b200.sampler import Cascade, CascadeConfig
    ms, sds = _models(use_cf)
    cfg = CascadeConfig(batch_size=2, num_surfaces=4, num_edges=3, use_cf=use_cf, class_label=6, schedule=schedule,
                        ddpm_steps=4, seed=3, decode=False)
    S = cfg.num_surfaces if use_cf else 2 * cfg.num_surfaces
    g = torch.Generator().manual_seed(9)
    init = {"surfPos": torch.randn(2, cfg.num_surfaces, 6, generator=g), "surfZ": torch.randn(2, S, 48, generator=g),
            "edgePos": torch.randn(2, S, 3, 6, generator=g), "edgeZV": torch.randn(2, S, 3, 18, generator=g)}
    bank = {}

    def step_noise(name, k, shape):
        key = (name, k)
        if key not in bank:
            bank[key] = torch.randn(tuple(shape), generator=g)
        return bank[key]

    ref = OC.run_cascade(sds, cfg, init, step_noise)
    out = Cascade(ms).run(cfg, init_noise=init, step_noise=step_noise)
    assert torch.equal(out["surfMask"].cpu(), ref["surfMask"])
    assert torch.equal(out["edgeM"].cpu(), ref["edgeM"])
    # compared on the valid slots: with mask-aware token compaction (the default) padded tokens are not computed at all (the
    # reference computes them as queries and discards them, sample.py:245,284,305-312)
    sv, ev = ~ref["surfMask"], ~ref["edgeM"]
    valid = {"surfPos": slice(None), "surfZ": sv, "edgePos": sv, "edge_z": ev, "edgeV": ev}
    for k in ("surfPos", "surfZ", "edgePos", "edge_z", "edgeV"):
        err = rel_l2(out[k].cpu()[valid[k]], ref[k][valid[k]])
        print(f"cascade cf={use_cf} {k} rel_l2={err:.3e}")
        assert err < 2e-3, (k, err)     # 4 chained steps: per-forward bar is 1e-3 (tests/test_gpu_denoisers.py)


def test_hybrid_schedule_runs_and_counts_steps():
    """the shipped PNDM/DDPM hybrid (sample.py:128-153,191-202,...): 158+250, 209, 158+250, 209 network evaluations"""
    from brepgen_b200.sampler import Cascade, CascadeConfig
    ms, _ = _models(False)
    counts = {}
    for k, m in ms.items():
        orig = m.forward

        def wrapped(*a, _k=k, _o=orig, **kw):
            counts[_k] = counts.get(_k, 0) + 1
            return _o(*a, **kw)
        m.forward = wrapped
    cfg = CascadeConfig(batch_size=1, num_surfaces=3, num_edges=2, schedule="reference", decode=False, graph="off")
    out = Cascade(ms).run(cfg)
    assert counts == {"surfpos": 408, "surfz": 209, "edgepos": 408, "edgez": 209}
    # with CUDA graphs the two 250-step DDPM tails are captured once (warm-up + capture = 2 python-level forwards each) and
    # replayed; the PNDM parts stay eager
    counts.clear()
    casc = Cascade(ms)
    out_g = casc.run(CascadeConfig(batch_size=1, num_surfaces=3, num_edges=2, schedule="reference", decode=False, graph="on"))
    assert counts == {"surfpos": 160, "surfz": 209, "edgepos": 160, "edgez": 209} and casc.last_graph_steps == 500
    assert out_g["surfPos"].shape == out["surfPos"].shape
    assert out["surfPos"].shape == (1, 6, 6) and out["edgeV"].shape == (1, 6, 2, 6)
    assert all(torch.isfinite(v.float()).all() for v in out.values())


def test_step_noise_is_reproducible_and_rank_dependent():
    """in-kernel Philox step noise: same (seed, rank, stage) -> identical x_{t-1}; another rank or stage -> different noise;
    a CPU generator is honoured like diffusers' randn_tensor does (sampled on the CPU, moved to the device)"""
    from brepgen_b200.schedulers import DDPMScheduler
    g = torch.Generator().manual_seed(3)
    x, eps = torch.randn(2, 30, 6, generator=g).cuda(), torch.randn(2, 30, 6, generator=g).cuda()

    def run(*key):
        s = DDPMScheduler(clip_sample=True, clip_sample_range=3)
        s.set_noise_seed(*key)
        y = s.step(eps, 500, x).prev_sample
        return torch.stack([y, s.step(eps, 499, y).prev_sample])

    a, b, c, d = run(11, 0, 2), run(11, 0, 2), run(11, 1, 2), run(11, 0, 3)
    assert torch.equal(a, b)
    assert not torch.allclose(a, c) and not torch.allclose(a, d) and not torch.allclose(c, d)
    # the two consecutive steps of one run draw different noise (the offset advances)
    s = DDPMScheduler(clip_sample=True, clip_sample_range=3)
    s.set_noise_seed(1)
    zero = torch.zeros_like(x)
    n1 = s.step(zero, 500, zero).prev_sample
    n2 = s.step(zero, 500, zero).prev_sample
    assert not torch.allclose(n1, n2) and abs(float(n1.std() / n2.std()) - 1) < 0.2
    cg = torch.Generator().manual_seed(9)
    y1 = DDPMScheduler().step(eps, 500, x, generator=cg).prev_sample
    cg.manual_seed(9)
    y2 = DDPMScheduler().step(eps, 500, x, generator=cg).prev_sample
    assert torch.equal(y1, y2)
