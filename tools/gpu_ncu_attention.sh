#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|attn_fwd" -s 2 -c 2 -o gpurun_out/r19_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r19_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r19_attn.ncu-rep
