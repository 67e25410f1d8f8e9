#!/bin/bash
# two MMA issuer warps + one-m-tile weight-gradient CTAs for small pixel counts: parity, A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_wgrad_gpu.py tests/test_network_gpu.py -m gpu -q -x > gpurun_out/r02_tests21.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests21.log
tail -3 gpurun_out/r02_tests21.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b21_new.log 2>&1
DMM_WGRAD_NA=4 $B > gpurun_out/r02_b21_na4.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_old.so DMM_WGRAD_NA=4 $B > gpurun_out/r02_b21_old.log 2>&1
$B > gpurun_out/r02_b21_new2.log 2>&1
DMM_WGRAD_NA1_PIXELS=2000000 $B > gpurun_out/r02_b21_na1all.log 2>&1
grep -h '"value"' gpurun_out/r02_b21_*.log | cut -c1-200
CASES="b1_conv2_fold b1_conv2_dgrad_k32 b1_conv2_dgrad_bnb b2_conv2_dgrad_bnb refine0 refine0_dgrad convT4_phase11 b1_conv1_k160_pro reduce4 b1_conv1_dgrad_n160"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for m in 0; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif6.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif6.log 2>&1
done
unset DMM_B200_LIB
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests21_all.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests21_all.log
tail -3 gpurun_out/r02_tests21_all.log
