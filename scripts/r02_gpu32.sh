#!/bin/bash
# half rings per MMA issuer warp: parity + A/B against the previous build
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests32.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests32.log
tail -3 gpurun_out/r02_tests32.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b32_new.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_prev.so $B > gpurun_out/r02_b32_prev.log 2>&1
$B > gpurun_out/r02_b32_new2.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_prev.so $B > gpurun_out/r02_b32_prev2.log 2>&1
grep -h '"value"' gpurun_out/r02_b32_*.log | cut -c1-200
CASES="b1_conv2_fold b1_conv2_dgrad_k32 refine0 refine0_dgrad convT4_phase11 convT1_phase11 reduce2 b1_conv1_k160_pro b1_conv1_dgrad_n160"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
DMM_IGEMM_PROF=1 timeout 150 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif12.log 2>&1
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif12.log | cut -c1-100
