#!/bin/bash
# ConvTranspose weight gradients grouped per output phase; BNB variant without spills; wgrad elected issue thread
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests26.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests26.log
tail -3 gpurun_out/r02_tests26.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b26_new.log 2>&1
DMM_WGRAD_GROUP_TAPS=0 $B > gpurun_out/r02_b26_nogroup.log 2>&1
DMM_CONV2_DGRAD_FUSED=1 $B > gpurun_out/r02_b26_conv2fused.log 2>&1
$B > gpurun_out/r02_b26_new2.log 2>&1
grep -h '"value"' gpurun_out/r02_b26_*.log | cut -c1-200
