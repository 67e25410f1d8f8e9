#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --profile-step"
export DMM_WGRAD_SIDE_STREAM=0
timeout 200 $B > gpurun_out/plain_bench3.log 2>&1 || exit 1
for k in "bn_relu_apply_kernel" "bn_bwd_maxpool_quad" "unfold_w7s2" "nchw_to_rows"; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -o gpurun_out/r02_ew_$k $B > gpurun_out/ncu_ew_$k.log 2>&1
done
ls -la gpurun_out/r02_ew_*.ncu-rep
