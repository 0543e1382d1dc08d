"""Experiment: how much does CUDA-graph replay of the whole fwd+bwd step gain over eager launches? (dropout off)"""
import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from lstc_vad_b200.harness import TrainStep, WORKLOADS, synthetic_step_inputs
from lstc_vad_b200 import functional as Fn
wl = WORKLOADS['ltn_sht']
dev = torch.device('cuda', 0)
step = TrainStep(wl, dev, seed=0, train_mode=False)
B = wl.batch_size
feats, labs = synthetic_step_inputs(wl, seed=1, device=dev)
def run():
    step.zero_grad()
    return step.forward_backward(feats, labs, B)
def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for _ in range(3): run()
print('eager ms/step', timeit(run))
# capture
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): run()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
step.zero_grad()
with torch.cuda.graph(g):
    terms = step.forward_backward(feats, labs, B)
print('captured; loss', terms['loss'].item())
print('graph ms/step', timeit(g.replay))
g.replay(); torch.cuda.synchronize()
print('loss after replay', terms['loss'].item())
print('eager again ms/step', timeit(run))
