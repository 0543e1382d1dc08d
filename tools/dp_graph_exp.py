"""Experiment (2 GPUs): does the data-parallel step (score all-gather + bucketed NCCL all-reduce on a side stream)
capture into a CUDA graph, and does it match the eager step?"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
from lstc_vad_b200.harness import TrainStep, GraphedTrainStep, WORKLOADS, synthetic_step_inputs
rank = int(os.environ['RANK']); torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
wl = WORKLOADS['ltn_sht']; B = wl.batch_size
step = TrainStep(wl, dev, seed=0, train_mode=True, process_group=dist.group.WORLD)
feats, labs = synthetic_step_inputs(wl, seed=10 + rank, device=dev)
def timeit(fn, n=8):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def eager():
    step.zero_grad(); step.forward_backward(feats, labs, B)
for _ in range(3): eager()
t_e = timeit(eager)
g = GraphedTrainStep(step, feats, labs, B, warmup=3)
t_g = timeit(lambda: g())
loss = g()['loss'].item()
if rank == 0: print(f'eager {t_e:.2f} ms  graph {t_g:.2f} ms  loss {loss:.4f}', flush=True)
g.close(); dist.destroy_process_group()
