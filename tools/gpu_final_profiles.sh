#!/bin/bash
# Round-end evidence on one B200: launch list (+ DRAM traffic) of one train step, ncu --set full of one launch of each
# attention kernel, the stand-alone kernel timings, compute-sanitizer over the attention unit tests, and the bench line.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list exit $?"
for spec in "fwd49 49 attn_fwd" "bwd49 49 attn_bwd" "fwd81 81 attn_fwd" "bwd81 81 attn_bwd" "bwd19 19 attn_bwd"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc" -s 3 -c 1 -o gpurun_out/r2_attn_$1 -f python tools/kernel_bench.py --reps 1 --L $2 --only "$3 dropout" > /dev/null 2>&1
  echo "ncu $1 exit $?"
done
for l in 49 81 19 17; do timeout 300 python tools/kernel_bench.py --L $l --only attn; done > gpurun_out/r2_kernel_bench_attn.txt 2>&1
timeout 300 python tools/kernel_bench.py --only ln >> gpurun_out/r2_kernel_bench_attn.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn_fwd or attn_bwd or attn_dropout" -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/r2_bench_default.log | cut -c1-300
ls -la gpurun_out/r2_attn_*.ncu-rep
