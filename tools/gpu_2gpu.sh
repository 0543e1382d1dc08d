#!/bin/bash
# 2-GPU checks: distributed tests, data-parallel bench (with strong-scaling / parity legs), sharded label sweep + round
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/g2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_label_sweep.py -m gpu -q -x -p no:cacheprovider > gpurun_out/g2_dist_test.log 2>&1
echo "dist tests exit $?" | tee gpurun_out/g2_summary.txt; tail -4 gpurun_out/g2_dist_test.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g2_bench_n2.log 2>&1; echo "bench n2 exit $?" | tee -a gpurun_out/g2_summary.txt
for f in gpurun_out/g2_bench_n2.log; do grep "^{" $f | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'launch')}, 'e2e', d['e2e']['value'], 'opt', d['with_optimizer'].get('ms_per_step'), d['with_optimizer'].get('launch'))
    print('strong', d.get('strong_scaling')); print('parity', d.get('dp_parity'))
except Exception as e:
    print('parse error', e)
"; done | tee -a gpurun_out/g2_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/label_sweep.py --videos 800 --round --steps-cap 4 > gpurun_out/g2_sweep.log 2>&1; echo "sweep exit $?" | tee -a gpurun_out/g2_summary.txt; tail -1 gpurun_out/g2_sweep.log | cut -c1-1500
