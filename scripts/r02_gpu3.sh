#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/r02_diag_trainer.py > gpurun_out/r02_diag_trainer.log 2>&1
python -m pytest tests/test_trainer_gpu.py tests/test_scatter_gpu.py tests/test_igemm_gpu.py tests/test_network_gpu.py -m gpu -q -s > gpurun_out/r02_tests3.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests3.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b_ns3.log 2>&1
DMM_IGEMM_NSLOT=1 DMM_IGEMM_SB1=6 $B > gpurun_out/r02_b_ns1.log 2>&1
DMM_IGEMM_NSLOT=2 $B > gpurun_out/r02_b_ns2.log 2>&1
DMM_IGEMM_NSLOT=3 DMM_IGEMM_SB1=2 $B > gpurun_out/r02_b_ns3_sb2.log 2>&1
for c in b1_conv1_k64_pro b2_conv1_k320_pro b3_conv1_k640_pro b1_conv1_dgrad_n160 b3_conv1_dgrad_n992; do
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c 2>&1 | tail -2 >> gpurun_out/r02_ig_prof3.log
done
python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r02_b_cfg5.log 2>&1
python bench.py --steps 5 --warmup 3 --workload cfg4 > gpurun_out/r02_b_cfg4b.log 2>&1
cat gpurun_out/r02_diag_trainer.log | tail -30; tail -3 gpurun_out/r02_tests3.log
