#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b39_base.log 2>&1
DMM_WGRAD_KPX=64 $B > gpurun_out/r02_b39_kpx64.log 2>&1
DMM_WGRAD_SPLIT_DIV=2 $B > gpurun_out/r02_b39_split2.log 2>&1
DMM_WGRAD_NA1_PIXELS=100000 $B > gpurun_out/r02_b39_na1_100k.log 2>&1
$B > gpurun_out/r02_b39_base2.log 2>&1
grep -h '"value"' gpurun_out/r02_b39_*.log | cut -c1-200
