#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "cuda_graph or dropout or attn_dropout" > gpurun_out/graph_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|Error|error" gpurun_out/graph_tests.log | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/graph_bench.log 2>&1; echo "bench exit $?"
tail -3 gpurun_out/graph_bench.log | cut -c1-300
