#!/bin/bash
# The measurements prepared at the end of round 1, one gpurun call (~3 min):
#   gpurun --timeout 600 -- 'bash tools/round2_first_run.sh > gpurun_out/round2_first_run.log 2>&1'
set -x
# tests added after the round-1 GPU budget ended (the driver one is xfail(strict=False) until it has passed once)
python -m pytest tests/test_gpu_zz_dedup_golden.py tests/test_gpu_zz_driver_golden.py tests/test_gpu_zz_config1_surfpos.py -q -rxX -s
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates tools/pipe_rates.cu && /tmp/pipe_rates
python tools/attn_check.py                                            # default kernel: parity + 716 TF/s at B = 64
BG_ATTN_PS=1 python tools/attn_check.py                               # persistent: 711
BG_ATTN_PS=1 BG_ATTN_TK=1 BG_ATTN_POLY=0 python tools/attn_check.py   # token-dense exponential section, MUFU only (unmeasured)
BG_ATTN_PS=1 BG_ATTN_TK=1 python tools/attn_check.py                  # ... with 25 % of the exponentials on the FMA pipe
BG_ATTN_PS=1 BG_ATTN_TK=1 BG_ATTN_PP=0 python tools/attn_check.py     # same code without the token
BG_ATTN_PS=1 BG_ATTN_PP=0 BG_ATTN_STAGGER=800 python tools/attn_check.py  # eager form, no token, warpgroup 1 starts 0.8 us late
BG_ATTN_PS=1 BG_ATTN_SPEC=1 BG_ATTN_PP=0 python tools/attn_check.py   # no row max, no token (with the token: 537)
BG_ATTN_PS=1 BG_ATTN_SPEC=1 BG_ATTN_PP=0 BG_ATTN_POLY=0 python tools/attn_check.py
python tools/gemm_time.py                                             # 2-CTA GEMMs as measured in round 1
BG_GEMM_PF=1 python tools/gemm_time.py                                # prefetching residual epilogue (unmeasured)
BG_GEMM_PF=2 python tools/gemm_time.py                                # TMA-staged residual epilogue (unmeasured, never run)
BG_GEMM_PF=2 python -m pytest tests/test_gpu_ops.py -q -k gemm         # ... its parity incl. ragged M and in-place residual
BG_GEMM_PF=2 python -m pytest tests/test_gpu_denoisers.py -q -k "vs_oracle"
