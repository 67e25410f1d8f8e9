#!/bin/bash
export PYTHONPATH=/root/repo
for dv in 1 2 4; do
  echo "== split_div $dv" >> gpurun_out/exp10_wg.log
  for c in b2_conv1_k320 b3_conv1_k640 b4_conv1_k768 reduce1 b3_conv2; do
    DMM_WGRAD_SPLIT_DIV=$dv DMM_WGRAD_PROF=1 timeout 300 python scripts/bench_wgrad.py $c 2>&1 | tail -2 >> gpurun_out/exp10_wg.log
  done
done
timeout 600 python -m pytest tests/test_elementwise_gpu.py -x -q > gpurun_out/exp10_test.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v31.json > gpurun_out/bench_v31.log 2>&1
