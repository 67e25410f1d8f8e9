#!/bin/bash
# fused tap issue (resident weights) + no divisions / clock reads in the MMA loop: parity, A/B, phase counters of the KxK cases
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_network_gpu.py -m gpu -q -x > gpurun_out/r02_tests18.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests18.log
tail -3 gpurun_out/r02_tests18.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b18_new.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_old.so $B > gpurun_out/r02_b18_old.log 2>&1
$B > gpurun_out/r02_b18_new2.log 2>&1
grep -h '"value"' gpurun_out/r02_b18_*.log | cut -c1-200
CASES="b1_conv2_fold b2_conv2 b1_conv2_dgrad_bnb b2_conv2_dgrad_bnb b3_conv2_dgrad_bnb b1_conv2_dgrad_k32 refine0 refine0_dgrad convT4_phase11 b3_conv1_k992_pro b4_conv1_k768_pro"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for m in 0 1 4 7 8 16 24 64 120; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif3.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif3.log 2>&1
done
