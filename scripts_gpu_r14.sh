#!/bin/bash
mkdir -p gpurun_out
export LSTC_GEMM_2CTA=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -p no:cacheprovider -k "gemm" > gpurun_out/r14_gemm2.log 2>&1
echo "2cta gemm tests exit $?" | tee gpurun_out/r14_summary.txt; grep -E "passed|failed|timed out|bad=[1-9]" gpurun_out/r14_gemm2.log | cut -c1-400 | tail -12
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r14_bench_2cta.log 2>&1; echo "bench 2cta exit $?" | tee -a gpurun_out/r14_summary.txt
tail -1 gpurun_out/r14_bench_2cta.log | cut -c1-200
unset LSTC_GEMM_2CTA
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r14_bench_1cta.log 2>&1; echo "bench 1cta exit $?" | tee -a gpurun_out/r14_summary.txt
tail -1 gpurun_out/r14_bench_1cta.log | cut -c1-200
