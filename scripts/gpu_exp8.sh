#!/bin/bash
export PYTHONPATH=/root/repo
timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/exp8_test.log 2>&1
for c in refine0 refine0_dgrad b1_conv2 b1_conv2_dgrad b1_conv1_k160_pro convT4_phase11; do
  DMM_IGEMM_PROF=1 timeout 300 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/exp8_ig.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v30.json > gpurun_out/bench_v30.log 2>&1
