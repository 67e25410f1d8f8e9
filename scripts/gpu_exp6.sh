#!/bin/bash
export PYTHONPATH=/root/repo
timeout 600 python -m pytest tests/test_igemm_gpu.py -x -q > gpurun_out/exp6_test_ig.log 2>&1
timeout 600 python -m pytest tests/test_network_gpu.py tests/test_elementwise_gpu.py -x -q > gpurun_out/exp6_test.log 2>&1
DMM_IGEMM_PROF=1 timeout 300 python scripts/bench_igemm.py refine1_fold > gpurun_out/exp6_ig.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v28.json > gpurun_out/bench_v28.log 2>&1
