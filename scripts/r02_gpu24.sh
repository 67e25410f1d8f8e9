#!/bin/bash
# one elected thread runs the MMA issue loop: parity + A/B against the previous build; sub-phases of the fused BN-backward statistics
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_network_gpu.py -m gpu -q -x > gpurun_out/r02_tests24.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests24.log
tail -3 gpurun_out/r02_tests24.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b24_new.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_prev.so $B > gpurun_out/r02_b24_prev.log 2>&1
DMM_CONV2_DGRAD_FUSED=0 $B > gpurun_out/r02_b24_new_unfused.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_prev.so DMM_CONV2_DGRAD_FUSED=0 $B > gpurun_out/r02_b24_prev_unfused.log 2>&1
grep -h '"value"' gpurun_out/r02_b24_*.log | cut -c1-200
CASES="b1_conv2_fold b1_conv2_dgrad_k32 b1_conv2_dgrad_bnb b2_conv2_dgrad_bnb refine0 refine0_dgrad convT4_phase11 b1_conv1_k160_pro b1_conv1_dgrad_n160"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
DMM_IGEMM_PROF=1 timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif9.log 2>&1
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif9.log | cut -c1-100
