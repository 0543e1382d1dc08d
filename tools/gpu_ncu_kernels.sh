#!/bin/bash
# ncu --set full of the non-GEMM kernels, driven by the micro-benchmark (few launches, small report)
mkdir -p gpurun_out
NAME=${1:-r31_kernels}
FILTER=${2:-"attn_bwd|attn_fwd|ln_fwd|ln_bwd_kernel"}
ONLY=${3:-"dropout"}
SKIP=${4:-6}
COUNT=${5:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$FILTER" -s $SKIP -c $COUNT -o gpurun_out/$NAME -f python tools/kernel_bench.py --reps 2 --only "$ONLY" > gpurun_out/${NAME}_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/$NAME.ncu-rep
