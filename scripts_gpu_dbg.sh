#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts_dbg_gemm.py > gpurun_out/dbg_gemm.log 2>&1
echo "dbg exit $?"; grep -v "bad=0/" gpurun_out/dbg_gemm.log | cut -c1-330 | head -30
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "layernorm" > gpurun_out/dbg_ln.log 2>&1; echo "ln tests exit $?"; tail -2 gpurun_out/dbg_ln.log
