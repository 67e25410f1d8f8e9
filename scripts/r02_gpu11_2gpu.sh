#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r02_tests_nccl.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests_nccl.log
grep -n "worst parameters" gpurun_out/r02_tests_nccl.log | cut -c1-1500
tail -3 gpurun_out/r02_tests_nccl.log
