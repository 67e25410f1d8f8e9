#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests14.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests14.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_fin1.log 2>&1
DMM_FUSE_FINALIZE=0 $B > gpurun_out/r02_b_fin0.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5_fold.log 2>&1
DMM_FOLD_EVAL_BN=0 timeout 400 python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5_nofold.log 2>&1
tail -3 gpurun_out/r02_tests14.log
