#!/bin/bash
set -x
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r02_b43_base.log 2>&1
DMM_WGRAD_SMALL_CTAS=74 $B > gpurun_out/r02_b43_s74.log 2>&1
DMM_WGRAD_SMALL_CTAS=96 $B > gpurun_out/r02_b43_s96.log 2>&1
DMM_WGRAD_SMALL_CTAS=111 $B > gpurun_out/r02_b43_s111.log 2>&1
DMM_WGRAD_SMALL_CTAS=74 DMM_WGRAD_SMALL_TILES=600 $B > gpurun_out/r02_b43_s74_600.log 2>&1
$B > gpurun_out/r02_b43_base2.log 2>&1
DMM_WGRAD_SMALL_CTAS=74 $B > gpurun_out/r02_b43_s74b.log 2>&1
