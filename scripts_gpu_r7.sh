#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r7_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/r7_summary.txt; tail -5 gpurun_out/r7_all.log
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_models.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r7_models.log 2>&1
echo "models+pipeline exit $?" | tee -a gpurun_out/r7_summary.txt; grep -E "passed|failed|ROC-AUC" gpurun_out/r7_models.log | tail -4
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r7_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r7_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r7_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
