#!/bin/bash
# Turns the files one run of tools/gpu_final_profiles.sh + tools/gpu_ncu_kernels.sh left in gpurun_out/ (prefix $1) into the
# tracked summaries under profiles/ (run from the repo root, no GPU needed).
P=${1:-r2f}
G=gpurun_out/$P
python profiles/summarize_launches.py ${G}_launches_traffic.csv > profiles/r2_step_launches_traffic.txt
{
  echo "# ncu --set full --clock-control none, one launch each (third launch of the process), tools/gpu_ncu_kernels.sh; LTN-SHT layer shapes (62,720 rows)"
  echo "# order: fwd ffn w1 (bias+relu) | fwd ffn w2 (bias+dropout+residual) | dgrad ffn w2 (relu mask + fused column sums) | dgrad ffn w1 | wgrad ffn w1 (split-K) | fwd out-proj (dropout+residual)"
  python profiles/summarize_ncu.py ${G}_ncu_gemm_*.csv
} > profiles/r2_ncu_gemm_summary.txt
{
  echo "# ncu --set full --clock-control none, one launch each, tools/gpu_ncu_kernels.sh: tcgen05 attention at L = 49 / 81 (dropout + rel-pos bias), LayerNorm at 62,720 x 2048"
  python profiles/summarize_ncu.py $(ls ${G}_ncu_attn_*.csv ${G}_ncu_ln_*.csv 2>/dev/null || ls gpurun_out/r2m_ncu_attn_*.csv gpurun_out/r2m_ncu_ln_*.csv)
} > profiles/r2_ncu_attention_summary.txt
{
  echo "# tools/kernel_bench.py --L {49,81,19,17} --only attn ; --only ln  (one B200, CUDA events, L2 flushed between repetitions; 1280 windows, 8 heads x 256; HBM peak = MEASURED_PEAKS.json hbm_gbs 6544 GB/s)"
  echo "# blocks in order: L=49 (LTN-SHT), L=81 (UBnormal), L=19 (UCF), L=17 (STN), LayerNorm at 62,720 x 2048"
  cat ${G}_kernel_bench.txt
  echo "# tools/gemm_bench.py --reps 20 (each case alone, back to back; run-to-run drift of several per cent: compare with the A/B list below)"
  cat ${G}_gemm_bench.txt
  if [ -f gpurun_out/gx_ab.log ]; then echo "# tools/gemm_ab.py (variants alternated inside one process; run before the eight-warp epilogue)"; cat gpurun_out/gx_ab.log; fi
  if [ -f gpurun_out/e2e_probe.log ]; then echo "# tools/e2e_probe.py 20"; grep -v Warning gpurun_out/e2e_probe.log | grep -v "return Variable"; fi
} > profiles/r2_kernel_bench.txt
{
  echo "# bench.py lines of the run (one B200): default (headline), --impl reference, other workloads"
  for f in ${G}_bench_default.log ${G}_bench_reference.log ${G}_wl_ltn_ubnormal.log ${G}_wl_ltn_ucf.log ${G}_wl_stn_sht.log; do grep "^{" $f; done
} > profiles/r2_bench_lines.jsonl
ls -la profiles/r2_*
