#!/bin/bash
set -x
mkdir -p gpurun_out
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
CASES="convT1_phase11 convT2_phase11 convT3_phase11 reduce1 reduce2 reduce3 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b4_conv1_dgrad_n768 b2_conv1_dgrad_n320"
for nt in 256 128; do
  echo "=== NTILE_MAX $nt" >> gpurun_out/r02_whatif14.log
  DMM_NTILE_MAX=$nt DMM_IGEMM_PROF=1 timeout 150 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif14.log 2>&1
done
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif14.log | cut -c1-100
unset DMM_B200_LIB
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b35_nt256.log 2>&1
DMM_NTILE_MAX=128 $B > gpurun_out/r02_b35_nt128.log 2>&1
grep -h '"value"' gpurun_out/r02_b35_*.log | cut -c1-200
