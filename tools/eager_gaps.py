"""Where an eagerly launched train step idles: kernel timeline of two steps (torch.profiler / CUPTI), gaps between
consecutive kernels, and the launches that precede the largest gaps."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstc_vad_b200.harness import WORKLOADS, TrainStep, synthetic_step_inputs  # noqa: E402

dev = torch.device("cuda:0")
wl = WORKLOADS["ltn_sht"]
step = TrainStep(wl, dev, seed=0)
feats, labs = synthetic_step_inputs(wl, seed=1, device=dev)
for _ in range(4):
    step.zero_grad()
    step.forward_backward(feats, labs, wl.batch_size)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step.zero_grad()
        step.forward_backward(feats, labs, wl.batch_size)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"kernels {len(evs)}  span {(t1 - t0) / 1e3:.2f} ms  busy {busy / 1e3:.2f} ms  idle {(t1 - t0 - busy) / 1e3:.2f} ms  (3 steps)")
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    gaps.append((g, a.name[:60], b.name[:60], (a.time_range.end - a.time_range.start)))
gaps.sort(reverse=True)
import collections
by = collections.Counter()
for g, a, b, d in gaps:
    if g > 0:
        by[(a.split("<")[0][-40:], b.split("<")[0][-40:])] += g
print("largest single gaps (us): gap | after kernel (dur us) -> before kernel")
for g, a, b, d in gaps[:15]:
    print(f"  {g:8.1f} | {a} ({d:.0f}) -> {b}")
print("gap totals by kernel pair (us over 3 steps):")
for k, v in by.most_common(15):
    print(f"  {v:9.1f}  {k[0]} -> {k[1]}")
small = sum(g for g, *_ in gaps if 0 < g <= 20)
print(f"sum of gaps <= 20 us: {small / 1e3:.2f} ms; > 20 us: {sum(g for g, *_ in gaps if g > 20) / 1e3:.2f} ms")
