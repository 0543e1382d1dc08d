#!/bin/bash
# quick loop: kernel tests of the touched kernels + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "layernorm or casts or colsum" > gpurun_out/quick_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/quick_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/quick_bench.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/quick_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
