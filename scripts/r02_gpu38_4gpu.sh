#!/bin/bash
# 4 GPUs: does capping NCCL's CTAs (fewer SMs taken from the persistent compute kernels) help the weak-scaling step time?
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
NCCL_MAX_CTAS=8 timeout 300 $R --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02b_scale_4gpu_cta8.log 2>&1
NCCL_MAX_CTAS=4 timeout 300 $R --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02b_scale_4gpu_cta4.log 2>&1
timeout 300 $R --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02b_scale_4gpu_b.log 2>&1
NCCL_MAX_CTAS=2 timeout 300 $R --master-port 29524 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02b_scale_4gpu_cta2.log 2>&1
grep -h '"value"' gpurun_out/r02b_scale_4gpu_*.log | cut -c1-220
