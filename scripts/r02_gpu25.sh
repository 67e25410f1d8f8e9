#!/bin/bash
# fused BN-backward coefficients from shared memory; which data gradients should carry the fused reduce?
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests25.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests25.log
tail -3 gpurun_out/r02_tests25.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b25_default.log 2>&1
DMM_CONV2_DGRAD_FUSED=1 $B > gpurun_out/r02_b25_conv2fused.log 2>&1
DMM_FUSE_BN_BWD_REDUCE=0 $B > gpurun_out/r02_b25_nofuse.log 2>&1
$B > gpurun_out/r02_b25_default2.log 2>&1
grep -h '"value"' gpurun_out/r02_b25_*.log | cut -c1-200
CASES="b1_conv2_dgrad_bnb b2_conv2_dgrad_bnb"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
DMM_IGEMM_PROF=1 timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif10.log 2>&1
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif10.log | cut -c1-100
