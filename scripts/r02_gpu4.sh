#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_strict_gpu.py tests/test_trainer_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/r02_tests4.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests4.log
grep -n "strict tf32\|tf32 conv\|passed\|failed\|B=32\|golden\]" gpurun_out/r02_tests4.log | tail -30
