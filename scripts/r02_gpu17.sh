#!/bin/bash
# fast tile decode + vectorised statistics: parity, A/B against the previous build, phase counters
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_wgrad_gpu.py tests/test_network_gpu.py -m gpu -q -x > gpurun_out/r02_tests17.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests17.log
tail -3 gpurun_out/r02_tests17.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b17_new.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_old.so $B > gpurun_out/r02_b17_old.log 2>&1
$B > gpurun_out/r02_b17_new2.log 2>&1
DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_old.so $B > gpurun_out/r02_b17_old2.log 2>&1
grep -h '"value"' gpurun_out/r02_b17_*.log | cut -c1-200
CASES="b1_conv1_k64_pro b1_conv1_k160_pro b2_conv1_k320_pro b3_conv1_k640_pro b1_conv1_dgrad_n160 b2_conv1_dgrad_n320 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b1_conv2_fold b1_conv2_dgrad reduce4"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for m in 0 1 4 7 8 16 24 64 120; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif2.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif2.log 2>&1
done
unset DMM_B200_LIB
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests17_all.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests17_all.log
tail -3 gpurun_out/r02_tests17_all.log
