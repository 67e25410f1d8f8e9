#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_elementwise_gpu.py tests/test_network_gpu.py tests/test_igemm_gpu.py -m gpu -q -x > gpurun_out/r02_tests47.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests47.log
tail -3 gpurun_out/r02_tests47.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b47_new.log 2>&1
$B > gpurun_out/r02_b47_new2.log 2>&1
