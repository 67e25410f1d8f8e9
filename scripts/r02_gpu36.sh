#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b36_base.log 2>&1
DMM_HEAD_FOLD0=1 $B > gpurun_out/r02_b36_fold0.log 2>&1
grep -h '"value"' gpurun_out/r02_b36_*.log | cut -c1-200
