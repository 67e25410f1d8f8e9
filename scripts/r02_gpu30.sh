#!/bin/bash
set -x
mkdir -p gpurun_out
export DMM_B200_LIB=$PWD/dmmfods_b200/libdmmfods_b200_whatif.so
CASES="convT1_phase11 convT2_phase11 convT3_phase11 reduce1 reduce2 reduce3 refine0 refine0_dgrad"
for m in 0 1 4 8; do
  echo "=== whatif mask $m" >> gpurun_out/r02_whatif11.log
  DMM_IGEMM_PROF=1 DMM_IGEMM_WHATIF=$m timeout 150 python scripts/bench_igemm.py $CASES >> gpurun_out/r02_whatif11.log 2>&1
done
grep -h "^convT\|^reduce\|^refine\|===" gpurun_out/r02_whatif11.log | cut -c1-100
