#!/bin/bash
# 8 GPUs, final build of round 2: the driver's scaling command
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02b_scale_8gpu.log 2>&1
tail -c 300 gpurun_out/r02b_scale_8gpu.log
