#!/bin/bash
export PYTHONPATH=/root/repo
for cfg in "DMM_IGEMM_MIN_SB=2" "DMM_IGEMM_MIN_SB=3" "DMM_IGEMM_MIN_SB=4" "DMM_IGEMM_MSUB=2" "DMM_IGEMM_MSUB=2 DMM_IGEMM_MIN_SB=4"; do
  echo "== $cfg" >> gpurun_out/exp9_ig.log
  for c in refine0 refine0_dgrad b1_conv2 b1_conv2_dgrad convT4_phase11 b2_conv2; do
    env $cfg DMM_IGEMM_PROF=1 timeout 300 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/exp9_ig.log
  done
done
