#!/bin/bash
export PYTHONPATH=/root/repo
DMM_FUSE_BN_BWD_REDUCE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v27_nofuse.json > gpurun_out/bench_v27_nofuse.log 2>&1
DMM_FUSE_BN_PROLOGUE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-step --dump-ops gpurun_out/ops_v27_nopro.json > gpurun_out/bench_v27_nopro.log 2>&1
