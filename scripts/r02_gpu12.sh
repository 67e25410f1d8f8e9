#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests12.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests12.log
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_nospill.log 2>&1
$B > gpurun_out/r02_b_nospill2.log 2>&1
DMM_IGEMM_WRES=0 DMM_DGRAD_PACK32=0 $B > gpurun_out/r02_b_nospill_wres0.log 2>&1
timeout 200 python bench.py --steps 5 --warmup 3 --workload cfg4 > gpurun_out/r02_b_cfg4c.log 2>&1
tail -3 gpurun_out/r02_tests12.log
