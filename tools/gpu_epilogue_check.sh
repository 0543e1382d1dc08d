mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/e8
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -x -p no:cacheprovider > ${O}_tests.log 2>&1; echo "tests exit $?"; tail -3 ${O}_tests.log
timeout 300 python tools/gemm_bench.py --reps 20 > ${O}_gemm_bench.txt 2>&1; cat ${O}_gemm_bench.txt
i=0
for c in "fwd out-proj drop" "dgrad ffn w2" "fwd ffn w2" "fwd ffn w1"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none -k regex:"gemm_bf16" -s 2 -c 1 -o /tmp/ncu/gemm_$i -f python tools/gemm_bench.py --reps 1 --only "$c" > /dev/null 2>&1
  ncu -i /tmp/ncu/gemm_$i.ncu-rep --page raw --csv > ${O}_ncu_gemm_$i.csv 2>/dev/null
done
python profiles/summarize_ncu.py ${O}_ncu_gemm_*.csv | grep gemm_bf16 | cut -d'|' -f1,2,5,6
for r in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_bench_$r.log 2>&1; python - ${O}_bench_$r.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), "windows/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"]), "gemm frac", round(d["roofline"]["frac"], 3), {k: round(v["tflops"]) for k, v in d["roofline"]["by_operand_layout"].items()}, d["clocks"]["sm_mhz"], {k: (round(v["frac"], 2), round(v["ms"], 3), round(v["ms_right_after_train_steps"], 3)) for k, v in d["roofline_hbm"].items()})
PY
done
