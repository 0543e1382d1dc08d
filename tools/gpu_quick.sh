#!/bin/bash
# kernel + model tests, then two bench passes (A/B of the LSTC_* env knob given as $1, e.g. LSTC_DGRAD_TRANSPOSED_W=0)
mkdir -p gpurun_out
O=gpurun_out/q
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -x -p no:cacheprovider > ${O}_tests.log 2>&1; echo "tests exit $?" | tee ${O}_summary.txt; tail -4 ${O}_tests.log
if [ -n "$1" ]; then env $1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_bench_a.log 2>&1; echo "bench A ($1) exit $?" | tee -a ${O}_summary.txt; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_bench_b.log 2>&1; echo "bench B exit $?" | tee -a ${O}_summary.txt
if [ -n "$1" ]; then env $1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_bench_a2.log 2>&1; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_bench_b2.log 2>&1
for f in a b a2 b2; do [ -f ${O}_bench_$f.log ] && python - ${O}_bench_$f.log $f >> ${O}_summary.txt <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[2], round(d["value"]), "windows/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"]), " eager", round(d["eager"]["ms_per_step"], 2), "opt", round(d["with_optimizer"]["ms_per_step"], 2), "gemm frac", round(d["roofline"]["frac"], 3), {k: round(v["tflops"]) for k, v in d["roofline"]["by_operand_layout"].items()}, d["clocks"]["sm_mhz"])
PY
done
cat ${O}_summary.txt
