#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests9.log 2>&1; echo "rc $?" >> gpurun_out/r02_tests9.log
for c in b1_conv1_k64_pro b2_conv1_k320_pro b1_conv1_dgrad_n160 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 reduce4; do
  DMM_IGEMM_PROF=1 timeout 120 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/r02_ig_prof9.log
done
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_stl1.log 2>&1
DMM_IGEMM_ST_LANE=0 $B > gpurun_out/r02_b_stl0.log 2>&1
DMM_IGEMM_ST_LANE=1 DMM_IGEMM_NSLOT=3 $B > gpurun_out/r02_b_stl1_ns3.log 2>&1
tail -3 gpurun_out/r02_tests9.log
