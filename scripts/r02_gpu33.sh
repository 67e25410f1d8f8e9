#!/bin/bash
set -x
mkdir -p gpurun_out
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
for ms in 0 1 2 4; do
  echo "=== MSUB $ms" >> gpurun_out/r02_whatif13.log
  DMM_IGEMM_MSUB=$ms DMM_IGEMM_PROF=1 timeout 150 python scripts/bench_igemm.py refine0 refine0_dgrad convT4_phase11 b1_conv2_dgrad_k32 >> gpurun_out/r02_whatif13.log 2>&1
done
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif13.log | cut -c1-100
grep "ig2" gpurun_out/r02_whatif13.log | awk 'NR%6==0' | cut -c1-330
