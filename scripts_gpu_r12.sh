#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r12_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/r12_summary.txt; grep -E "^FAILED|^E  .*Assert|passed|failed" gpurun_out/r12_all.log | cut -c1-300 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r12_summary.txt
tail -1 gpurun_out/r12_bench.log | cut -c1-300
