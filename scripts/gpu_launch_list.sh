#!/bin/bash
# launch list only (first half of gpu_profile.sh): gpu time + DRAM bytes of every launch of one training step
set -x
TAG=${1:-r01}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --profile-step --dump-ops gpurun_out/${TAG}_ops.json"
export DMM_WGRAD_SIDE_STREAM=0      # serialised launches: the launch order is the op order
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
