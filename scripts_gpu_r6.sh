#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -p no:cacheprovider -k "layernorm" > gpurun_out/r6_ln.log 2>&1
echo "ln kernels exit $?" | tee gpurun_out/r6_summary.txt; grep -E "passed|failed" gpurun_out/r6_ln.log | tail -2
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r6_models.log 2>&1
echo "models exit $?" | tee -a gpurun_out/r6_summary.txt; grep -E "passed|failed" gpurun_out/r6_models.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/r6_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r6_summary.txt; grep smoke: gpurun_out/r6_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r6_summary.txt
timeout 900 ncu --set full --clock-control none -k regex:"attn_fwd|attn_bwd|ln_fwd_kernel|ln_bwd_kernel" -s 12 -c 8 -o gpurun_out/r6_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/r6_summary.txt
ls -la gpurun_out/
