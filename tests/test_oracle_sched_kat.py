"""CPU tests: the scheduler oracle against the known answers of diffusers' OWN scheduler tests.

diffusers (requirements.txt:5 of the reference pins 0.27) is absent from the image, so its code cannot be run here; its
test-suite, however, publishes known answers for full sampling loops over a deterministic dummy model
(tests/schedulers/test_scheduler_ddpm.py::test_full_loop_no_noise, test_scheduler_pndm.py::test_full_loop_no_noise and
::test_full_loop_with_no_set_alpha_to_one).  The loops, the dummy model / sample and the expected |x| sum / mean below are
restated from those tests (the three sum / mean pairs are mutually consistent: sum / 768 = mean).  Reproducing them pins
the beta schedule, the timestep tables, the DDPM posterior mean / fixed_small variance / clipping / noise injection and the
PNDM Runge-Kutta + linear-multistep tables and transfer formula of oracle/schedulers.py to diffusers' arithmetic.
Tolerances are diffusers' own (1e-2 on the sum, 1e-3 on the mean)."""
import torch

from oracle.schedulers import DDPMOracle, PNDMOracle

N_ELEMS = 4 * 3 * 8 * 8


def dummy_sample_deter():
    return (torch.arange(N_ELEMS).reshape(3, 8, 8, 4) / N_ELEMS).permute(3, 0, 1, 2).contiguous()


def dummy_model(sample, t):
    t = float(t)
    return sample * t / (t + 1)


DDPM_KAT = (258.9606, 0.3372)           # DDPMScheduler(clip_sample=True) over all 1000 steps, noise from torch.manual_seed(0)
PNDM_KAT = (198.1318, 0.2580)           # PNDMScheduler(), set_timesteps(10), PRK then PLMS steps
PNDM_B01_KAT = (186.9482, 0.2434)       # ... with beta_start = 0.01


def test_ddpm_full_loop_matches_diffusers_known_answer():
    sch = DDPMOracle(clip_sample=True, clip_sample_range=1.0)
    g = torch.manual_seed(0)
    x = dummy_sample_deter()
    for t in reversed(range(1000)):
        eps = dummy_model(x, t)
        noise = torch.randn(eps.shape, generator=g) if t > 0 else None    # diffusers draws noise only for t > 0
        x = sch.step(eps, t, x, noise)
    assert abs(float(x.abs().sum()) - DDPM_KAT[0]) < 1e-2
    assert abs(float(x.abs().mean()) - DDPM_KAT[1]) < 1e-3


def _pndm_loop(**kw):
    sch = PNDMOracle(**kw)
    sch.set_timesteps(10)
    x = dummy_sample_deter()
    for t in sch.timesteps:
        x = sch.step(dummy_model(x, int(t)), int(t), x)
    return x


def test_pndm_full_loop_matches_diffusers_known_answer():
    x = _pndm_loop()
    assert abs(float(x.abs().sum()) - PNDM_KAT[0]) < 1e-2
    assert abs(float(x.abs().mean()) - PNDM_KAT[1]) < 1e-3


def test_pndm_beta_start_variant_matches_diffusers_known_answer():
    x = _pndm_loop(beta_start=0.01)
    assert abs(float(x.abs().sum()) - PNDM_B01_KAT[0]) < 1e-2
    assert abs(float(x.abs().mean()) - PNDM_B01_KAT[1]) < 1e-3
