#!/bin/bash
# round-end evidence: launch list (+DRAM traffic) of one step, and ncu --set full of one launch of each hot kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/final_ncu_bench.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|attn_fwd|ln_fwd|ln_bwd_kernel" -s 3 -c 1 -o gpurun_out/final_attn_fwd -f python tools/kernel_bench.py --reps 1 --only "attn_fwd dropout" > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd" -s 3 -c 1 -o gpurun_out/final_attn_bwd -f python tools/kernel_bench.py --reps 1 --only "attn_bwd dropout" > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd" -s 4 -c 1 -o gpurun_out/final_ln_fwd -f python tools/kernel_bench.py --reps 1 --only "ln_fwd bf16->bf16" > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_kernel" -s 3 -c 1 -o gpurun_out/final_ln_bwd -f python tools/kernel_bench.py --reps 1 --only "ln_bwd dropout+dxsum" > /dev/null 2>&1
ls -la gpurun_out/final_*.ncu-rep
