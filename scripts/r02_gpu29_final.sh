#!/bin/bash
# final evidence of round 2 (1 GPU): tests, smoke, the bench lines the driver will ask for, other workloads / APIs
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r02_final2_tests.log 2>&1; echo "rc $?" >> gpurun_out/r02_final2_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r02_final2_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final2_bench_20_5.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_final2_bench_20_5b.log 2>&1
timeout 300 python bench.py > gpurun_out/r02_final2_bench_default.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_final2_bench_reference.log 2>&1
timeout 300 python bench.py --api module --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final2_bench_module.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 > gpurun_out/r02_final2_bench_cfg4.log 2>&1
grep -n "passed\|failed" gpurun_out/r02_final2_tests.log; tail -3 gpurun_out/r02_final2_smoke.log
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02_final2_bench_cfg5.log 2>&1
