#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cpu_overhead_exp.py > gpurun_out/cpu_overhead.log 2>&1; echo "exit $?"; head -60 gpurun_out/cpu_overhead.log | cut -c1-200
