#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b42_base.log 2>&1
DMM_WGRAD_SMALL_CTAS=48 $B > gpurun_out/r02_b42_s48.log 2>&1
DMM_WGRAD_SMALL_CTAS=74 $B > gpurun_out/r02_b42_s74.log 2>&1
DMM_WGRAD_SMALL_CTAS=48 DMM_WGRAD_SMALL_TILES=2400 $B > gpurun_out/r02_b42_s48_2400.log 2>&1
$B > gpurun_out/r02_b42_base2.log 2>&1
grep -h '"value"' gpurun_out/r02_b42_*.log | cut -c1-200
