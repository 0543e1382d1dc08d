#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/r18_bench.log | cut -c1-200
timeout 600 python scripts_graph_exp.py > gpurun_out/r18_graph.log 2>&1; echo "graph exp exit $?"; tail -8 gpurun_out/r18_graph.log
