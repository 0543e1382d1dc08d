#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r4_models.log 2>&1
echo "models exit $?" | tee gpurun_out/r4_summary.txt
grep -E "passed|failed" gpurun_out/r4_models.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/r4_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r4_summary.txt; grep smoke: gpurun_out/r4_smoke.log
# full ncu capture of the three GEMM layouts and the attention / LN kernels (3 launches each, after warm-up)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|attn_fwd|attn_bwd|ln_fwd_kernel|ln_bwd_kernel" -s 150 -c 40 -o gpurun_out/r4_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r4_ncu.log 2>&1
echo "ncu full exit $?" | tee -a gpurun_out/r4_summary.txt
ls -la gpurun_out/r4_prof.ncu-rep
