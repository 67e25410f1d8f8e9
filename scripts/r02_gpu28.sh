#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b28_new.log 2>&1
DMM_FUSE_MIN_PIXELS=0 $B > gpurun_out/r02_b28_fuseall.log 2>&1
$B > gpurun_out/r02_b28_new2.log 2>&1
DMM_FUSE_MIN_PIXELS=0 $B > gpurun_out/r02_b28_fuseall2.log 2>&1
grep -h '"value"' gpurun_out/r02_b28_*.log | cut -c1-200
timeout 600 python -m pytest tests/test_network_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02_tests28.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests28.log
tail -3 gpurun_out/r02_tests28.log
