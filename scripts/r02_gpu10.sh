#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/r02_diag_twopass.py > gpurun_out/r02_diag_twopass.log 2>&1
timeout 900 bash scripts/gpu_profile.sh r02 > gpurun_out/r02_profile.log 2>&1
cat gpurun_out/r02_diag_twopass.log
