#!/bin/bash
# which shared-memory descriptor mode of igemm v2 is right on this hardware?  (each mode in its own process)
mkdir -p gpurun_out
run() {
  echo "=== $1" | tee -a gpurun_out/ig2_modes.log
  env $2 timeout 600 python -m pytest tests/test_igemm_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -15 >> gpurun_out/ig2_modes.log
  echo "rc=$?" >> gpurun_out/ig2_modes.log
}
: > gpurun_out/ig2_modes.log
run "v2 dense pitch, base_offset 0" "X=1"
run "v2 pitch8, base_offset 0" "DMM_IGEMM_PITCH8=1"
run "v2 pitch8, base_offset from address" "DMM_IGEMM_PITCH8=1 DMM_IGEMM_DESC_BO=1"
run "v2 dense pitch, base_offset from address" "DMM_IGEMM_DESC_BO=1"
run "v1" "DMM_IGEMM_V1=1"
cat gpurun_out/ig2_modes.log
