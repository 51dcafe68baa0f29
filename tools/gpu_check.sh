#!/bin/bash
# one gpurun call: the GPU test suite + launch lists (ncu gpu__time_duration) of one forward at the configs[1] shape and at the
# edge-stage shape
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_surfpos_b64_s30.csv python tools/profile_forward.py --kind surfpos --batch 64 --surfaces 30 --iters 1 --time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_edgepos_b8.csv python tools/profile_forward.py --kind edgepos --batch 8 --iters 1 --time
timeout 120 python tools/profile_forward.py --kind surfpos --batch 64 --surfaces 30 --iters 50 --time
timeout 120 python tools/profile_forward.py --kind edgepos --batch 64 --iters 3 --time
