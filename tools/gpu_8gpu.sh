#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/g8_ngpus.txt
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g8_bench_n$n.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/g8_bench_n$n.log 2>&1
  fi
  echo "bench n$n exit $?" | tee -a gpurun_out/g8_summary.txt
  tail -1 gpurun_out/g8_bench_n$n.log | cut -c1-200
done
