#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r11_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/r11_summary.txt; tail -3 gpurun_out/r11_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r11_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r11_summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r11_bench_ref.log 2>&1
# per-launch time + DRAM traffic of every kernel of one step (traffic field of the roofline)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r11_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "ncu traffic exit $?" | tee -a gpurun_out/r11_summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r11_smi.txt
