#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/full_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/full_summary.txt; grep -E "^FAILED|^E  .*Assert|passed|failed" gpurun_out/full_all.log | cut -c1-300 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/full_summary.txt
tail -1 gpurun_out/full_bench.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/full_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
