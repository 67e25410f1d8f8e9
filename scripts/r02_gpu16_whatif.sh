#!/bin/bash
# what-if attribution of the 1x1 implicit-GEMM epilogue (experiment build with -DDMM_IGEMM_WHATIF -DDMM_IGEMM_PHASE_PROF) + A-ring cap A/B
set -x
mkdir -p gpurun_out
CASES="b1_conv1_k64_pro b1_conv1_k160_pro b2_conv1_k320_pro b1_conv1_dgrad_n160 b2_conv1_dgrad_n320 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992"
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for m in 0 1 2 3 4 7 8 16 24 32 64 96 104 128 129; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 120 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif.log 2>&1
done
unset DMM_B200_LIB
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_sa4.log 2>&1
DMM_IGEMM_SA_CAP1=2 $B > gpurun_out/r02_b_sa2.log 2>&1
DMM_IGEMM_SA_CAP1=3 $B > gpurun_out/r02_b_sa3.log 2>&1
$B > gpurun_out/r02_b_sa4b.log 2>&1
grep -h '"value"' gpurun_out/r02_b_sa*.log | cut -c1-200
