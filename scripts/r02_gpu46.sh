#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b46_base.log 2>&1
DMM_PDL=1 $B > gpurun_out/r02_b46_pdl.log 2>&1
$B > gpurun_out/r02_b46_base2.log 2>&1
DMM_PDL=1 $B > gpurun_out/r02_b46_pdl2.log 2>&1
