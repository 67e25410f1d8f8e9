#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b19_fused.log 2>&1
DMM_CONV2_DGRAD_FUSED=0 $B > gpurun_out/r02_b19_unfused.log 2>&1
$B > gpurun_out/r02_b19_fused2.log 2>&1
DMM_CONV2_DGRAD_FUSED=0 $B > gpurun_out/r02_b19_unfused2.log 2>&1
grep -h '"value"' gpurun_out/r02_b19_*.log | cut -c1-200
CASES="b1_conv2_fold b2_conv2 b1_conv2_dgrad_k32 refine0 refine0_dgrad convT4_phase11 b1_conv1_k160_pro reduce4"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for m in 0 4; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif4.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif4.log 2>&1
done
unset DMM_B200_LIB
for na in 4 2 1; do
  echo "=== DMM_WGRAD_NA $na" >> gpurun_out/r02_wgrad_na.log
  DMM_WGRAD_NA=$na timeout 120 python scripts/bench_wgrad.py b2_conv1_k512 b3_conv1_k640 b3_conv1_k1024 b4_conv1_k768 b4_conv1_k1024 reduce1 >> gpurun_out/r02_wgrad_na.log 2>&1
done
cat gpurun_out/r02_wgrad_na.log
