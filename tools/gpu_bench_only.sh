#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_only.log 2>&1; echo "bench exit $?"
tail -3 gpurun_out/bench_only.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_oracle_golden.py -q -p no:cacheprovider > gpurun_out/bench_only_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/bench_only_tests.log
