#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r15_all.log 2>&1
echo "all gpu tests exit $?" | tee gpurun_out/r15_summary.txt; grep -E "^FAILED|^E  .*Assert|passed|failed" gpurun_out/r15_all.log | cut -c1-300 | tail -8
LSTC_GEMM_2CTA=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k gemm > gpurun_out/r15_gemm1cta.log 2>&1; echo "1cta gemm tests exit $?" | tee -a gpurun_out/r15_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r15_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r15_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r15_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/r15_summary.txt
tail -1 gpurun_out/r15_bench.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r15_launches_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16_tcgen05_2cta" -s 30 -c 6 -o gpurun_out/r15_gemm2cta -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r15_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
