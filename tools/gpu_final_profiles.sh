#!/bin/bash
# Round-end evidence on one B200 (prefix $1, default r2f; $2 = "tests" also runs every GPU test first): the bench line of
# the headline and of the other workloads, the launch list (+ DRAM traffic) of one train step, stand-alone kernel timings.
# The ncu --set full captures are separate (tools/gpu_ncu_kernels.sh): gpurun brings back at most 64 MiB.
P=${1:-r2f}
mkdir -p gpurun_out
O=gpurun_out/$P
if [ "$2" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > ${O}_tests.log 2>&1; echo "gpu tests exit $?" | tee ${O}_summary.txt; tail -3 ${O}_tests.log | cut -c1-200
fi
timeout 900 python bench.py > ${O}_bench_default.log 2>&1; echo "bench exit $?" | tee -a ${O}_summary.txt; tail -1 ${O}_bench_default.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${O}_bench_reference.log 2>&1; echo "reference arm exit $?" | tee -a ${O}_summary.txt
for wl in ltn_ubnormal ltn_ucf stn_sht; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${O}_wl_$wl.log 2>&1; echo "bench $wl exit $?" | tee -a ${O}_summary.txt
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${O}_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extras > ${O}_ncu_bench.log 2>&1
echo "launch list exit $?" | tee -a ${O}_summary.txt
for l in 49 81 19 17; do timeout 300 python tools/kernel_bench.py --L $l --only attn; done > ${O}_kernel_bench.txt 2>&1
timeout 300 python tools/kernel_bench.py --only ln >> ${O}_kernel_bench.txt 2>&1
timeout 300 python tools/gemm_bench.py --reps 20 > ${O}_gemm_bench.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" -p no:cacheprovider > ${O}_sanitizer_gemm.log 2>&1
echo "memcheck gemm exit $?" | tee -a ${O}_summary.txt; tail -3 ${O}_sanitizer_gemm.log | cut -c1-200
cat ${O}_summary.txt
du -sh gpurun_out
