"""How long does the host need to ENQUEUE one eager train step (no sync) vs how long the GPU needs to run it?"""
import sys, time, torch, cProfile, pstats
sys.path.insert(0, '.')
from lstc_vad_b200.harness import TrainStep, WORKLOADS, synthetic_step_inputs
wl = WORKLOADS['ltn_sht']; dev = torch.device('cuda', 0); B = wl.batch_size
step = TrainStep(wl, dev, seed=0, train_mode=True)
feats, labs = synthetic_step_inputs(wl, seed=1, device=dev)
def run():
    step.zero_grad(); step.forward_backward(feats, labs, B)
for _ in range(3): run()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): run()
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'host enqueue {1e3*(t1-t0)/n:.2f} ms/step   gpu {e0.elapsed_time(e1)/n:.2f} ms/step   wall {1e3*(t2-t0)/n:.2f} ms/step')
# tiny batch: pure host cost per step
f2, l2 = synthetic_step_inputs(wl, seed=2, batch_size=1, device=dev)
def run_small():
    step.zero_grad(); step.forward_backward(f2, l2, 1)
for _ in range(3): run_small()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): run_small()
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f'B=1 step wall {1e3*(t1-t0)/20:.2f} ms (host-bound)')
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run_small()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
