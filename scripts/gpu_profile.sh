#!/bin/bash
# ncu evidence for profiles/: launch list + DRAM bytes of every launch of one full training step, full captures of the top kernels
set -x
TAG=${1:-r01}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --profile-step --dump-ops gpurun_out/${TAG}_ops.json"
export DMM_WGRAD_SIDE_STREAM=0      # serialised launches: the launch order is the op order
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
python scripts/bench_igemm.py refine0 > gpurun_out/plain_ig.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:igemm2 -s 3 -c 1 -o gpurun_out/${TAG}_igemm2_refine0 python scripts/bench_igemm.py refine0 > gpurun_out/ncu_ig.log 2>&1
python scripts/bench_igemm.py b1_conv1_k160 > gpurun_out/plain_ig1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:igemm2 -s 3 -c 1 -o gpurun_out/${TAG}_igemm2_conv1x1 python scripts/bench_igemm.py b1_conv1_k160 > gpurun_out/ncu_ig1.log 2>&1
python scripts/bench_wgrad.py b1_conv2 > gpurun_out/plain_wg.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -o gpurun_out/${TAG}_wgrad_b1conv2 python scripts/bench_wgrad.py b1_conv2 > gpurun_out/ncu_wg.log 2>&1
$B > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bn_bwd_contrib -s 5 -c 1 -o gpurun_out/${TAG}_bn_bwd_contrib $B > gpurun_out/ncu_bn.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:head_input_kernel -c 1 -o gpurun_out/${TAG}_head_input $B > gpurun_out/ncu_head.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grad_gather -s 3 -c 1 -o gpurun_out/${TAG}_grad_gather $B > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out/*.ncu-rep
# round 2: the two batched pre-processing kernels of BASELINE config 4 (one --set full capture each)
B4="python bench.py --workload cfg4 --steps 2 --warmup 3 --graph 0 --profile-step"
$B4 > gpurun_out/plain_bench_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lidar_splat_tile -c 1 -o gpurun_out/${TAG}_lidar_splat_tile $B4 > gpurun_out/ncu_splat.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:heatmap_tile -c 1 -o gpurun_out/${TAG}_heatmap_tile $B4 > gpurun_out/ncu_heat.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
