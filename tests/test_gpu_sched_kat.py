"""GPU tests: the PRODUCT schedulers (fused step kernels through the C ABI) run the full sampling loops of diffusers' own
scheduler tests and must reproduce their published known answers (see tests/test_oracle_sched_kat.py for the provenance)."""
import pytest
import torch

from test_oracle_sched_kat import DDPM_KAT, PNDM_B01_KAT, PNDM_KAT, dummy_model, dummy_sample_deter

pytestmark = pytest.mark.gpu


def test_ddpm_scheduler_full_loop_known_answer():
    from brepgen_b200.schedulers import DDPMScheduler
    sch = DDPMScheduler(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                        clip_sample=True, clip_sample_range=1.0)
    g = torch.manual_seed(0)               # CPU generator: noise is drawn on the CPU and copied, like diffusers' randn_tensor
    x = dummy_sample_deter().cuda()
    for t in reversed(range(1000)):
        x = sch.step(dummy_model(x, t), t, x, generator=g).prev_sample
    x = x.cpu()
    assert abs(float(x.abs().sum()) - DDPM_KAT[0]) < 1e-2
    assert abs(float(x.abs().mean()) - DDPM_KAT[1]) < 1e-3


@pytest.mark.parametrize("kw,kat", [({}, PNDM_KAT), ({"beta_start": 0.01}, PNDM_B01_KAT)])
def test_pndm_scheduler_full_loop_known_answer(kw, kat):
    from brepgen_b200.schedulers import PNDMScheduler
    sch = PNDMScheduler(num_train_timesteps=1000, beta_start=kw.get("beta_start", 0.0001), beta_end=0.02,
                        beta_schedule="linear")
    sch.set_timesteps(10)
    x = dummy_sample_deter().cuda()
    for t in sch.timesteps:
        x = sch.step(dummy_model(x, int(t)), t, x).prev_sample
    x = x.cpu()
    assert abs(float(x.abs().sum()) - kat[0]) < 1e-2
    assert abs(float(x.abs().mean()) - kat[1]) < 1e-3
