#!/bin/bash
export PYTHONPATH=/root/repo
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/exp14_test.log 2>&1
DMM_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v35_pdl0.log 2>&1
DMM_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v35_pdl1.log 2>&1
DMM_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v35_pdl0b.log 2>&1
DMM_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v35_pdl1b.log 2>&1
