#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/g2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -p no:cacheprovider > gpurun_out/g2_dist_test.log 2>&1
echo "dist test exit $?" | tee gpurun_out/g2_summary.txt; tail -5 gpurun_out/g2_dist_test.log
timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g2_bench_n1.log 2>&1; echo "bench n1 exit $?" | tee -a gpurun_out/g2_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/g2_bench_n2.log 2>&1; echo "bench n2 exit $?" | tee -a gpurun_out/g2_summary.txt
tail -2 gpurun_out/g2_bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/g2_bench_ref.log 2>&1; echo "bench ref n2 exit $?" | tee -a gpurun_out/g2_summary.txt
