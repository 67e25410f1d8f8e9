#!/bin/bash
# x/y role swap for KxK tiles on small images (igemm2 launcher, wgrad family plans): parity + A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests31.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests31.log
tail -3 gpurun_out/r02_tests31.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b31_new.log 2>&1
DMM_IGEMM_SWAP_XY=0 DMM_WGRAD_SWAP_XY=0 $B > gpurun_out/r02_b31_noswap.log 2>&1
$B > gpurun_out/r02_b31_new2.log 2>&1
DMM_IGEMM_SWAP_XY=0 DMM_WGRAD_SWAP_XY=0 $B > gpurun_out/r02_b31_noswap2.log 2>&1
grep -h '"value"' gpurun_out/r02_b31_*.log | cut -c1-200
