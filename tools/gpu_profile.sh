#!/bin/bash
# ncu evidence for the round: (1) launch list of the bench command (per-kernel share of the step), (2) one --set full capture
# of the dominant kernel (edge-stage attention at B = 256, L = 4000) for the DRAM traffic and pipe utilisation
set -x
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_bench.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -o gpurun_out/r02_attn_b256 -f \
  env B=256 python tools/attn_check.py > gpurun_out/r02_attn_ncu.out 2>&1
ls -la gpurun_out/*.ncu-rep
