#!/bin/bash
# round 2, GPU call 1: full -m gpu suite, A/B of the balanced pack kernels and the TMA L2 promotion, launch list, 1x1 probes
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_tests1.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests1.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_default.log 2>&1
DMM_BALANCED_PACK=0 $B > gpurun_out/r02_b_pack0.log 2>&1
DMM_TMA_L2_PROMO=128 $B > gpurun_out/r02_b_promo128.log 2>&1
DMM_TMA_L2_PROMO=0 $B > gpurun_out/r02_b_promo0.log 2>&1
for c in b1_conv1_k64_pro b1_conv1_k160_pro b2_conv1_k320_pro b3_conv1_k640_pro b3_conv1_k992_pro b4_conv1_k768_pro b1_conv1_dgrad_n160 b2_conv1_dgrad_n320 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b4_conv1_dgrad_n768; do
  python scripts/bench_igemm.py $c >> gpurun_out/r02_ig_plain.log 2>&1
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c >> gpurun_out/r02_ig_prof.log 2>&1
done
python scripts/bench_wgrad.py b1_conv1_k160 b2_conv1_k320 b3_conv1_k640 b4_conv1_k768 > gpurun_out/r02_wg_plain.log 2>&1
bash scripts/gpu_launch_list.sh r02a > gpurun_out/r02_launch_list.log 2>&1
tail -3 gpurun_out/r02_tests1.log; tail -c 600 gpurun_out/r02_b_default.log
