#!/bin/bash
export PYTHONPATH=/root/repo
timeout 900 python -m pytest tests/test_elementwise_gpu.py tests/test_wgrad_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/exp19_test.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v38.json > gpurun_out/bench_v38.log 2>&1
