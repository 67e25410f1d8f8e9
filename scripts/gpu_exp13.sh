#!/bin/bash
export PYTHONPATH=/root/repo
for al in 8 16; do
DMM_COL_ALIGN=$al timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v34_$al.json > gpurun_out/bench_v34_$al.log 2>&1
done
timeout 600 python -m pytest tests/test_network_gpu.py -x -q > gpurun_out/exp13_test.log 2>&1
