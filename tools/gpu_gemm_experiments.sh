#!/bin/bash
# GEMM variants on one B200: bf16 outputs through TMA stores vs register stores; input gradients against the row-major
# weight (MN-major B operand) vs a transposed copy (K-major B).
mkdir -p gpurun_out
O=gpurun_out/gx
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm -p no:cacheprovider > ${O}_tests.log 2>&1; echo "gemm tests exit $?" | tee ${O}_summary.txt; tail -3 ${O}_tests.log
timeout 300 python tools/gemm_bench.py --check > ${O}_check.log 2>&1; echo "check exit $?" | tee -a ${O}_summary.txt; grep -c OK ${O}_check.log; grep MISMATCH ${O}_check.log
LSTC_GEMM_TMA_STORE=0 timeout 300 python tools/gemm_bench.py --reps 20 > ${O}_time_direct.log 2>&1
timeout 300 python tools/gemm_bench.py --reps 20 > ${O}_time_tma.log 2>&1
paste <(cut -c1-50 ${O}_time_direct.log) <(cut -c31-50 ${O}_time_tma.log) | tee -a ${O}_summary.txt
