#!/bin/bash
# first GPU bring-up: per-kernel parity, each group in its own process (a trap kills the CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "gemm_plain" "gemm and not gemm_plain" "layernorm" "attn_fwd" "attn_bwd or attn_dropout" "not gemm and not layernorm and not attn"; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$grp" -s -p no:cacheprovider > "gpurun_out/r1_${tag}.log" 2>&1
  echo "group [$grp] exit $?" | tee -a gpurun_out/r1_summary.txt
  tail -5 "gpurun_out/r1_${tag}.log"
done
