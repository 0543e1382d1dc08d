"""Where does the end-to-end loop lose time against the device-resident loop?  (one B200)

Replays the LTN-SHT train-step graphs like bench.py's e2e pass and switches the parts of the loop on and off:
H2D copy of the next batch on a side stream, D2H read of the loss + host wait for the previous step's loss.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lstc_vad_b200.harness import WORKLOADS, GraphedTrainStep, TrainStep, synthetic_step_inputs  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    wl = WORKLOADS["ltn_sht"]
    B = wl.batch_size
    step = TrainStep(wl, dev, seed=0, train_mode=True)
    host = [synthetic_step_inputs(wl, seed=100 + i, batch_size=B, pin=True) for i in range(2)]
    resident = [(f.to(dev), l.to(dev)) for f, l in host]
    for i in range(3):
        step.zero_grad()
        step.forward_backward(*resident[i % 2], B)
    torch.cuda.synchronize()
    graphs = [GraphedTrainStep(step, f, l, B, warmup=2) for f, l in resident]
    for g in graphs:
        g()
    torch.cuda.synchronize()
    copy_stream = torch.cuda.Stream()
    loss_host = [torch.empty(1, pin_memory=True) for _ in range(2)]

    def loop(h2d, d2h, first_outside=False, chunks=1):
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_read = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                if h2d:
                    f_dev, f_host = graphs[s].static_feats, host[s][0]
                    n = f_dev.shape[0]
                    for c in range(chunks):
                        a, b = c * n // chunks, (c + 1) * n // chunks
                        f_dev[a:b].copy_(f_host[a:b], non_blocking=True)
                    graphs[s].static_labs.copy_(host[s][1], non_blocking=True)
                ready[s].record(copy_stream)

        for s in range(2):
            consumed[s].record()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if first_outside:
            prefetch(0)
            torch.cuda.synchronize()
        w0 = time.perf_counter()
        t0.record()
        if not first_outside:
            prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            graphs[s].graph.replay()
            consumed[s].record()
            if d2h:
                loss_host[s].copy_(graphs[s].terms["loss"].detach().reshape(1), non_blocking=True)
                loss_read[s].record()
                if i >= 1:
                    loss_read[1 - s].synchronize()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / steps, (time.perf_counter() - w0) * 1e3 / steps

    def plain():
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            graphs[i % 2]()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / steps

    print(f"# {steps} steps per measurement, ms per step (CUDA events; wall clock in brackets)")
    for rep in range(2):
        print(f"resident graph loop                         {plain():7.3f}")
        for name, kw in (("events only (no H2D, no D2H)", dict(h2d=False, d2h=False)),
                         ("D2H loss + host wait, no H2D", dict(h2d=False, d2h=True)),
                         ("H2D next batch, no D2H", dict(h2d=True, d2h=False)),
                         ("H2D + D2H (bench e2e)", dict(h2d=True, d2h=True)),
                         ("H2D + D2H, H2D in 4 chunks", dict(h2d=True, d2h=True, chunks=4)),
                         ("H2D + D2H, first copy outside the timing", dict(h2d=True, d2h=True, first_outside=True))):
            ev, wall = loop(**kw)
            print(f"{name:44s}{ev:7.3f}  [{wall:7.3f}]", flush=True)
    for g in graphs:
        g.close()


if __name__ == "__main__":
    main()
