#!/bin/bash
export PYTHONPATH=/root/repo
DMM_TEST_FOLD64=1 timeout 600 python -m pytest tests/test_igemm_gpu.py -x -q -k "growth_folded" > gpurun_out/exp23_unit.log 2>&1
DMM_HEAD_FOLD0=1 timeout 900 python -m pytest tests/test_network_gpu.py -x -q > gpurun_out/exp23_test.log 2>&1
DMM_HEAD_FOLD0=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v41_0.json > gpurun_out/bench_v41_0.log 2>&1
DMM_HEAD_FOLD0=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v41_1.json > gpurun_out/bench_v41_1.log 2>&1
