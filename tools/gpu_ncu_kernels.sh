#!/bin/bash
# ncu --set full of one launch of each GEMM flavour of the FFN (forward w1 / w2, relu-mask dgrad, dgrad, wgrad) and of the
# attention kernels at L = 49; the reports are turned into raw-metric CSVs on the box (gpurun brings back <= 64 MiB) and
# summarised here with profiles/summarize_ncu.py --csv.
P=${1:-r2f}
WHAT=${2:-all}   # "gemm": only the GEMM captures
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/$P
i=0
for c in "fwd ffn w1" "fwd ffn w2" "dgrad ffn w2" "dgrad ffn w1" "wgrad ffn w1" "fwd out-proj drop"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none -k regex:"gemm_bf16" -s 2 -c 1 -o /tmp/ncu/gemm_$i -f python tools/gemm_bench.py --reps 1 --only "$c" > /dev/null 2>&1
  echo "ncu gemm '$c' exit $?"
  ncu -i /tmp/ncu/gemm_$i.ncu-rep --page raw --csv > ${O}_ncu_gemm_$i.csv 2>/dev/null
done
if [ "$WHAT" = "gemm" ]; then ls -la ${O}_ncu_*.csv; exit 0; fi
for spec in "fwd49 49 attn_fwd" "bwd49 49 attn_bwd" "fwd81 81 attn_fwd" "bwd81 81 attn_bwd"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:"attn_tc" -s 3 -c 1 -o /tmp/ncu/attn_$1 -f python tools/kernel_bench.py --reps 1 --L $2 --only "$3 dropout" > /dev/null 2>&1
  echo "ncu $1 exit $?"
  ncu -i /tmp/ncu/attn_$1.ncu-rep --page raw --csv > ${O}_ncu_attn_$1.csv 2>/dev/null
done
for k in "ln_fwd bf16->bf16" "ln_bwd dropout+dxsum"; do
  n=$(echo $k | cut -c1-6)
  timeout 300 ncu --set full --clock-control none -k regex:"ln_" -s 3 -c 1 -o /tmp/ncu/$n -f python tools/kernel_bench.py --reps 1 --only "$k" > /dev/null 2>&1
  echo "ncu $n exit $?"
  ncu -i /tmp/ncu/$n.ncu-rep --page raw --csv > ${O}_ncu_$n.csv 2>/dev/null
done
ls -la ${O}_ncu_*.csv; du -sh gpurun_out
