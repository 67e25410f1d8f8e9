#!/bin/bash
# ncu evidence for profiles/: launch list of one full training step + full captures of the dominant kernels
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --profile-step"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1400 --csv --log-file gpurun_out/r01_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
python scripts/bench_igemm.py refine0 > gpurun_out/plain_ig.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:igemm2 -s 3 -c 1 -o gpurun_out/r01_igemm2_refine0 python scripts/bench_igemm.py refine0 > gpurun_out/ncu_ig.log 2>&1
python scripts/bench_wgrad.py b1_conv2 > gpurun_out/plain_wg.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -o gpurun_out/r01_wgrad_b1conv2 python scripts/bench_wgrad.py b1_conv2 > gpurun_out/ncu_wg.log 2>&1
$B > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bn_bwd_apply_fast -s 40 -c 1 -o gpurun_out/r01_bn_bwd_apply $B > gpurun_out/ncu_bn.log 2>&1
ls -la gpurun_out/*.ncu-rep
