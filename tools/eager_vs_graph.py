"""Steady-state step time of the same train step launched eagerly vs replayed as a CUDA graph (no clock sampler)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstc_vad_b200.harness import WORKLOADS, GraphedTrainStep, TrainStep, synthetic_step_inputs  # noqa: E402

dev = torch.device("cuda:0")
wl = WORKLOADS["ltn_sht"]
step = TrainStep(wl, dev, seed=0)
feats, labs = synthetic_step_inputs(wl, seed=1, device=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    cpu = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, cpu


def eager():
    step.zero_grad()
    step.forward_backward(feats, labs, wl.batch_size)


for _ in range(5):
    eager()
print("eager  #1  %.2f ms/step (cpu launch %.2f ms/step)" % timed(eager, K), flush=True)
g = GraphedTrainStep(step, feats, labs, wl.batch_size, warmup=2)
g()
print("graph  #1  %.2f ms/step (cpu %.2f)" % timed(g, K), flush=True)
g.close()
print("eager  #2  %.2f ms/step (cpu launch %.2f ms/step)" % timed(eager, K), flush=True)
_lib_reset = g
g2 = GraphedTrainStep(step, feats, labs, wl.batch_size, warmup=2)
g2()
print("graph  #2  %.2f ms/step (cpu %.2f)" % timed(g2, K), flush=True)
