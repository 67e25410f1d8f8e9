#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02_*.ncu-rep
timeout 1500 bash scripts/gpu_profile.sh r02 > gpurun_out/r02_profile_run.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep gpurun_out/r02_launches.csv gpurun_out/r02_ops.json
