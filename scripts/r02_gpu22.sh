#!/bin/bash
# vectorised fused BN-backward statistics, switchable second MMA warp: parity + A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_network_gpu.py tests/test_elementwise_gpu.py -m gpu -q -x > gpurun_out/r02_tests22.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests22.log
tail -3 gpurun_out/r02_tests22.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b22_mma2.log 2>&1
DMM_IGEMM_MMA2=0 $B > gpurun_out/r02_b22_mma1.log 2>&1
$B > gpurun_out/r02_b22_mma2b.log 2>&1
DMM_IGEMM_MMA2=0 $B > gpurun_out/r02_b22_mma1b.log 2>&1
DMM_FUSE_BN_PROLOGUE_KXK=1 $B > gpurun_out/r02_b22_kxk.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_old.so DMM_WGRAD_NA=4 $B > gpurun_out/r02_b22_old.log 2>&1
grep -h '"value"' gpurun_out/r02_b22_*.log | cut -c1-200
CASES="b1_conv2_fold b1_conv2_dgrad_k32 b1_conv2_dgrad_bnb b2_conv2_dgrad_bnb b3_conv2_dgrad_bnb refine0 refine0_dgrad convT4_phase11"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for v in 1 0; do
  echo "=== MMA2 $v" >> gpurun_out/r02_whatif7.log
  DMM_IGEMM_MMA2=$v DMM_IGEMM_PROF=1 timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif7.log 2>&1
done
echo "=== MSUB 2" >> gpurun_out/r02_whatif7.log
DMM_IGEMM_MSUB=2 DMM_IGEMM_PROF=1 timeout 120 python scripts/bench_igemm.py refine0 refine0_dgrad >> gpurun_out/r02_whatif7.log 2>&1
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif7.log | cut -c1-100
