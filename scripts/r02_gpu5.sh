#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_tests5.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests5.log
DMM_IGEMM_WRES=0 DMM_DGRAD_PACK32=0 python -m pytest tests/test_igemm_gpu.py tests/test_network_gpu.py -m gpu -q > gpurun_out/r02_tests5b.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests5b.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B --dump-ops gpurun_out/r02_ops_wres1.json > gpurun_out/r02_b_wres1.log 2>&1
DMM_IGEMM_WRES=0 DMM_DGRAD_PACK32=0 $B --dump-ops gpurun_out/r02_ops_wres0.json > gpurun_out/r02_b_wres0.log 2>&1
DMM_IGEMM_WRES=1 DMM_DGRAD_PACK32=0 $B > gpurun_out/r02_b_wres1_p0.log 2>&1
for c in b1_conv1_k64_pro b1_conv1_k160_pro b2_conv1_k320_pro b3_conv1_k640_pro b1_conv1_dgrad_n160 b2_conv1_dgrad_n320 b3_conv1_dgrad_n640 b3_conv1_dgrad_n992 b1_conv2_fold b1_conv2_dgrad; do
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/r02_ig_prof5.log
done
grep -n "passed\|failed" gpurun_out/r02_tests5.log gpurun_out/r02_tests5b.log
