#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -p no:cacheprovider -k "attn" > gpurun_out/r5_attn.log 2>&1
echo "attn kernels exit $?" | tee gpurun_out/r5_summary.txt; grep -E "passed|failed" gpurun_out/r5_attn.log | tail -2
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r5_models.log 2>&1
echo "models exit $?" | tee -a gpurun_out/r5_summary.txt; grep -E "passed|failed" gpurun_out/r5_models.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/r5_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r5_summary.txt; grep smoke: gpurun_out/r5_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r5_summary.txt; tail -2 gpurun_out/r5_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
