#!/bin/bash
# 2 GPUs, final build of round 2: NCCL correctness test + scaling pair + the reference arm under torchrun; every command under its own timeout
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r02b_tests_nccl.log 2>&1; echo "rc $?" >> gpurun_out/r02b_tests_nccl.log
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_scale_1gpu.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_scale_2gpu.log 2>&1
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_scale_1gpu_b.log 2>&1
tail -5 gpurun_out/r02b_tests_nccl.log; tail -c 400 gpurun_out/r02b_scale_2gpu.log
