#!/bin/bash
# experiment batch: contrib kernel variants, grid caps, refine1 cycle counters
export PYTHONPATH=/root/repo
python -m pytest tests/test_elementwise_gpu.py -x -q -k deferred > gpurun_out/exp1_test.log 2>&1
for v in 0 1 2 4 5; do
  echo "== variant $v" >> gpurun_out/exp1_bn.log
  DMM_CONTRIB_VARIANT=$v python scripts/bench_bn.py 10 2>&1 | grep contrib >> gpurun_out/exp1_bn.log
done
for b in 2 3 4 6; do
  echo "== bps $b (variant 0 / 4)" >> gpurun_out/exp1_bn.log
  DMM_EW_BPS=$b python scripts/bench_bn.py 10 >> gpurun_out/exp1_bn.log 2>&1
  DMM_EW_BPS=$b DMM_CONTRIB_VARIANT=4 python scripts/bench_bn.py 10 2>&1 | grep contrib >> gpurun_out/exp1_bn.log
done
for c in refine1 refine1_dgrad refine0 refine0_dgrad; do
  DMM_IGEMM_PROF=1 python scripts/bench_igemm.py $c > gpurun_out/exp1_ig_$c.log 2>&1
done
