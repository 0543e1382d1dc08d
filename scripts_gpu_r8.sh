#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r8_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/r8_summary.txt; tail -8 gpurun_out/r8_all.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r8_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
