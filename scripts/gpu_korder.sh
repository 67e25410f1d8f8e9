for k in 0 1; do echo "== KORDER=$k"; DMM_IGEMM_KORDER=$k python scripts/bench_igemm.py; done 2>&1 | tee gpurun_out/korder.log
