#!/bin/bash
set -x
mkdir -p gpurun_out
for c in b1_conv1_k64_pro b2_conv1_k320_pro b3_conv1_k640_pro b1_conv1_dgrad_n160 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b1_conv2_dgrad reduce4; do
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/r02_ig_prof6.log
done
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_al64.log 2>&1
DMM_DA1_ALIGN=8 $B > gpurun_out/r02_b_al8.log 2>&1
python -m pytest tests/test_network_gpu.py tests/test_fullsize_gpu.py tests/test_strict_gpu.py -m gpu -q > gpurun_out/r02_tests6.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests6.log
tail -3 gpurun_out/r02_tests6.log
