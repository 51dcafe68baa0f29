#!/bin/bash
# ncu --set full of the encoder GEMMs (one launch per shape of one layer at M = 256 000 through tools/gemm_time.py)
set -x
ONCE=1 timeout 900 ncu --set full --clock-control none -k regex:gemm2 -c 5 -o gpurun_out/r02_gemm_b64 -f python tools/gemm_time.py > gpurun_out/r02_gemm_ncu.out 2>&1
tail -8 gpurun_out/r02_gemm_ncu.out; ls -la gpurun_out/
timeout 600 python -m pytest tests/test_gpu_post.py tests/test_gpu_vae.py -q 2>&1 | tail -4
