#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for cfg in "overlap layer" "tail layer" "tail single"; do
  set -- $cfg
  LSTC_DP_REDUCE=$1 LSTC_DP_BUCKETS=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/dp_${1}_${2}_n$N.log 2>&1
  tail -1 gpurun_out/dp_${1}_${2}_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1 $2 N=$N', round(d['ms_per_step'], 3), 'ms/step', round(d['value']), 'e2e', round(d['e2e']['value']), 'eager', round(d['eager']['ms_per_step'], 2))"
done
