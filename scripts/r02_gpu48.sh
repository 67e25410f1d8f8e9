#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b48_base.log 2>&1
DMM_IGEMM_SB1=4 $B > gpurun_out/r02_b48_sb4.log 2>&1
DMM_IGEMM_SB1=3 $B > gpurun_out/r02_b48_sb3.log 2>&1
$B > gpurun_out/r02_b48_base2.log 2>&1
DMM_IGEMM_SB1=4 $B > gpurun_out/r02_b48_sb4b.log 2>&1
