#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -k "ucf" > gpurun_out/r17_ucf.log 2>&1; echo "ucf pooling tests exit $?"; tail -2 gpurun_out/r17_ucf.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/r17_bench.log | cut -c1-200
