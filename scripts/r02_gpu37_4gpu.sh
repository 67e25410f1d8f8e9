#!/bin/bash
# 4 GPUs, final build of round 2: the driver's scaling command (every command under its own timeout)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02b_scale_4gpu.log 2>&1
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_scale_1gpu_c.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/r02b_ref_4gpu.log 2>&1
tail -c 300 gpurun_out/r02b_scale_4gpu.log; tail -c 300 gpurun_out/r02b_ref_4gpu.log
