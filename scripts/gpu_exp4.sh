#!/bin/bash
export PYTHONPATH=/root/repo
python -m pytest tests/test_elementwise_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/exp4_test.log 2>&1
DMM_IGEMM_PROF=1 python scripts/bench_igemm.py refine1_dgrad_v > gpurun_out/exp4_ig.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v27.json > gpurun_out/bench_v27.log 2>&1
DMM_HEAD_DGRAD_FUSED=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v27_nf.json > gpurun_out/bench_v27_nf.log 2>&1
