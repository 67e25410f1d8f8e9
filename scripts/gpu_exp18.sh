#!/bin/bash
export PYTHONPATH=/root/repo
timeout 900 python -m pytest tests/test_elementwise_gpu.py -x -q -k "unfold" > gpurun_out/exp18_test.log 2>&1
DMM_STEM_UNFOLD_MIN_C=9 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v37_0.json > gpurun_out/bench_v37_0.log 2>&1
DMM_STEM_UNFOLD_MIN_C=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v37_1.json > gpurun_out/bench_v37_1.log 2>&1
DMM_STEM_UNFOLD_MIN_C=9 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v37_0b.log 2>&1
DMM_STEM_UNFOLD_MIN_C=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v37_1b.log 2>&1
