#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_models.log 2>&1
echo "models exit $?" | tee gpurun_out/r2_summary.txt
tail -30 gpurun_out/r2_models.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r2_summary.txt; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r2_summary.txt; tail -5 gpurun_out/r2_bench.log
