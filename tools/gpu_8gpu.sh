#!/bin/bash
# One 8-GPU box: weak-scaling bench at N = 1, 2, 4, 8 (with the strong-scaling and data-parallel parity legs), the sharded
# label sweep over the 10,000-video corpus at N = 1 and 8, and one data-parallel co-teaching round at N = 8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l > gpurun_out/g8_ngpus.txt
run_bench() {  # $1 = N
  if [ "$1" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/g8_bench_n1.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29720 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/g8_bench_n$1.log 2>&1
  fi
  echo "bench n$1 exit $?" | tee -a gpurun_out/g8_summary.txt
}
: > gpurun_out/g8_summary.txt
for n in 1 2 4 8; do run_bench $n; done
timeout 600 python tools/label_sweep.py --videos 10000 > gpurun_out/g8_sweep_n1.log 2>&1; echo "sweep n1 exit $?" | tee -a gpurun_out/g8_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 tools/label_sweep.py --videos 10000 --round --out gpurun_out/g8_labels.npy > gpurun_out/g8_sweep_n8.log 2>&1; echo "sweep n8 exit $?" | tee -a gpurun_out/g8_summary.txt
rm -f gpurun_out/g8_labels.npy
python - <<'PY' | tee -a gpurun_out/g8_summary.txt
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/g8_bench_n{n}.log").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "parse error", e); continue
    base = base or d["value"]
    print(f"N={n}: {d['value']:.0f} windows/s ({d['ms_per_step']:.2f} ms/step, weak eff {d['value'] / (n * base):.3f}), e2e {d['e2e']['value']:.0f}, "
          f"with optimizer {d['with_optimizer']['ms_per_step']:.2f} ms ({d['with_optimizer'].get('launch')}), strong {d.get('strong_scaling', {}).get('ms_per_step')}, "
          f"parity {d.get('dp_parity', {}).get('grad_rel_l2_max')}")
for f in ("g8_sweep_n1", "g8_sweep_n8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("n_gpus", "value", "windows", "sweep_seconds", "sweep_wall_seconds_incl_merge_and_save")}, d.get("round"))
    except Exception as e:
        print(f, "parse error", e)
PY
