#!/bin/bash
# compute-sanitizer over the hand-written kernels' unit tests (memcheck: global/shared OOB incl. the TMA ring and
# the ragged LayerNorm groups; racecheck: shared-memory hazards of the attention stage ring and LayerNorm scratch)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 300 $SAN --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "attn_fwd or attn_bwd or layernorm or colsum or cls_prepend or head_tail or mil_loss" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/san_memcheck.log | tail -5
timeout 200 $SAN --tool memcheck --error-exitcode 7 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "gemm" > gpurun_out/san_memcheck_gemm.log 2>&1
echo "memcheck gemm exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_gemm.log | tail -3
timeout 200 $SAN --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "(attn_fwd or attn_bwd) and (3-49 or 2-81 or 5-19) or layernorm_bwd_dxsum" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/san_racecheck.log | tail -5
