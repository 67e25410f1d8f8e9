#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b41_base.log 2>&1
DMM_WGRAD_SIDE_STREAM=0 $B > gpurun_out/r02_b41_noside.log 2>&1
$B > gpurun_out/r02_b41_base2.log 2>&1
DMM_WGRAD_SIDE_STREAM=0 $B > gpurun_out/r02_b41_noside2.log 2>&1
grep -h '"value"' gpurun_out/r02_b41_*.log | cut -c1-200
