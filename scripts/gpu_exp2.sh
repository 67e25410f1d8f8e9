#!/bin/bash
export PYTHONPATH=/root/repo
python -m pytest tests/test_elementwise_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/exp2_test.log 2>&1
for b in 2 3 4; do
  echo "== gather bps $b" >> gpurun_out/exp2_bn.log
  DMM_GATHER_BPS=$b python scripts/bench_bn.py 10 >> gpurun_out/exp2_bn.log 2>&1
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v25.json > gpurun_out/bench_v25.log 2>&1
DMM_HEAD_BPS=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v25_h4.json > gpurun_out/bench_v25_h4.log 2>&1
DMM_HEAD_BPS=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v25_h2.json > gpurun_out/bench_v25_h2.log 2>&1
