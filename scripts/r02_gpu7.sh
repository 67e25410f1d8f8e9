#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests7.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests7.log
for c in b1_conv1_k64_pro b2_conv1_k320_pro b1_conv1_dgrad_n160 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b1_conv2_dgrad reduce4; do
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/r02_ig_prof7.log
done
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_lsu1.log 2>&1
DMM_IGEMM_LSU_STORE=0 $B > gpurun_out/r02_b_lsu0.log 2>&1
DMM_DA1_ALIGN=8 DMM_IGEMM_LSU_STORE=0 $B > gpurun_out/r02_b_lsu0_al8.log 2>&1
python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5_fold.log 2>&1
DMM_FOLD_EVAL_BN=0 python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5_nofold.log 2>&1
tail -3 gpurun_out/r02_tests7.log
