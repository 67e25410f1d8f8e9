#!/bin/bash
# round 2, GPU call 2: full -m gpu suite (no -x), store-engine micro-benchmark, new bench workloads / APIs
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_tests2.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests2.log
./scripts/ubench/store_rate > gpurun_out/r02_store_rate.log 2>&1
python bench.py --steps 5 --warmup 3 --api module --no-cpu-baseline > gpurun_out/r02_b_module.log 2>&1
python bench.py --steps 5 --warmup 3 --workload cfg4 > gpurun_out/r02_b_cfg4.log 2>&1
python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5.log 2>&1
DMM_TMA_L2_PROMO=128 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_b_full128.log 2>&1
tail -5 gpurun_out/r02_tests2.log; cat gpurun_out/r02_store_rate.log; tail -c 400 gpurun_out/r02_b_module.log; tail -c 400 gpurun_out/r02_b_cfg4.log; tail -c 600 gpurun_out/r02_b_cfg5.log
