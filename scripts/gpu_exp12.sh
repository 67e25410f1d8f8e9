#!/bin/bash
export PYTHONPATH=/root/repo
for al in 8 16 32; do
DMM_HEAD_LD_ALIGN=$al timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v33_$al.json > gpurun_out/bench_v33_$al.log 2>&1
done
timeout 600 python -m pytest tests/test_network_gpu.py -x -q > gpurun_out/exp12_test.log 2>&1
