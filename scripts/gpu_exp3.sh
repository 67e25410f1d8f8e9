#!/bin/bash
export PYTHONPATH=/root/repo
python -m pytest tests/test_elementwise_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/exp3_test.log 2>&1
python scripts/bench_bn.py 10 > gpurun_out/exp3_bn.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v26.json > gpurun_out/bench_v26.log 2>&1
