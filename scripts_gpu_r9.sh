#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "layernorm or adagrad or 64w" > gpurun_out/r9_tests.log 2>&1
echo "tests exit $?" | tee gpurun_out/r9_summary.txt; tail -3 gpurun_out/r9_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|attn_bwd" -s 4 -c 2 -o gpurun_out/r9_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/r9_summary.txt
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16" -s 30 -c 9 -o gpurun_out/r9_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r9_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
