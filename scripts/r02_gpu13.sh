#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests13.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests13.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_f1.log 2>&1
DMM_HEAD_MINB=3 $B > gpurun_out/r02_b_f1_head3.log 2>&1
DMM_IGEMM_WRES=0 DMM_DGRAD_PACK32=0 DMM_BALANCED_PACK=0 $B > gpurun_out/r02_b_f1_r1like.log 2>&1
$B > gpurun_out/r02_b_f1b.log 2>&1
tail -3 gpurun_out/r02_tests13.log
