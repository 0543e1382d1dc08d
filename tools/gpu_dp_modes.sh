#!/bin/bash
# data-parallel reduce experiments at N GPUs (default 2): overlap vs tail launch, NCCL CTA limits
N=${1:-2}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/dp_$tag.log 2>&1
  echo "$tag exit $? $(tail -1 gpurun_out/dp_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))' 2>/dev/null)"
}
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/dp_n1.log 2>&1
tail -1 gpurun_out/dp_n1.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("n1", round(d["value"]), d["ms_per_step"], d["eager"])'
run overlap LSTC_DP_REDUCE=overlap
run tail LSTC_DP_REDUCE=tail
run overlap_cta8 LSTC_DP_REDUCE=overlap NCCL_MAX_CTAS=8
run overlap_cta4 LSTC_DP_REDUCE=overlap NCCL_MAX_CTAS=4
run tail_cta32 LSTC_DP_REDUCE=tail NCCL_MIN_CTAS=32
