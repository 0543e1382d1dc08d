#!/bin/bash
mkdir -p gpurun_out
for wl in ltn_ucf ltn_ubnormal stn_sht; do
  timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_$wl.log 2>&1; echo "bench $wl exit $?"
  tail -1 gpurun_out/r20_bench_$wl.log | cut -c1-250
done
