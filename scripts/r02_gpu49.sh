#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
for i in 1 2; do
$B > gpurun_out/r02_b49_base_$i.log 2>&1
DMM_IGEMM_SB1=3 $B > gpurun_out/r02_b49_sb3_$i.log 2>&1
DMM_IGEMM_SB1=2 $B > gpurun_out/r02_b49_sb2_$i.log 2>&1
done
