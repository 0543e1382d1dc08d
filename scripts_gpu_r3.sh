#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1
echo "ncu launches exit $?" | tee gpurun_out/r3_summary.txt
wc -l gpurun_out/r3_launches.csv
