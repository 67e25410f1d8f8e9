#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b40_base.log 2>&1
DMM_IGEMM_MIN_SB=4 $B > gpurun_out/r02_b40_minsb4.log 2>&1
DMM_IGEMM_MIN_SB=3 $B > gpurun_out/r02_b40_minsb3.log 2>&1
grep -h '"value"' gpurun_out/r02_b40_*.log | cut -c1-200
for v in 2 4; do
echo "=== MIN_SB $v" >> gpurun_out/r02_whatif15.log
DMM_IGEMM_MIN_SB=$v timeout 150 python scripts/bench_igemm.py refine0 refine0_dgrad convT4_phase11 convT1_phase11 reduce2 >> gpurun_out/r02_whatif15.log 2>&1
done
grep -h "^b[0-9]\|^refine\|^convT\|^reduce\|===" gpurun_out/r02_whatif15.log | cut -c1-100
