"""Multi-GPU parity (needs >= 2 GPUs, NCCL): the bag-sharded data-parallel train step (score all-gather, global
MIL loss, bucketed gradient all-reduce overlapped with backward) reproduces the single-GPU step on the same
global batch — loss terms and every parameter gradient (dropout off)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _workload():
    from lstc_vad_b200.harness import Workload
    return Workload("dp_test", 256, 512, 3, 16, 4, 4, n_layers=2, n_head=2, d_k=64, dropouts=(0, 0, 0, 0))


def _global_batch(wl, device):
    from lstc_vad_b200.harness import synthetic_step_inputs
    return synthetic_step_inputs(wl, seed=3, batch_size=wl.batch_size, device=device)


def _worker(rank, world, port, out, grad_dtype):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if grad_dtype == "fp32":  # also covers the per-layer buckets launched from the backward hooks
        os.environ.update(LSTC_DP_BUCKETS="layer", LSTC_DP_REDUCE="overlap")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from lstc_vad_b200.harness import TrainStep
    wl = _workload()
    feats, labs = _global_batch(wl, dev)
    B, P = wl.batch_size, wl.part_num
    Bl = B // world
    nh = B * P
    sel = torch.cat([torch.arange(rank * Bl * P, (rank + 1) * Bl * P), nh + torch.arange(rank * Bl * P, (rank + 1) * Bl * P)]).to(dev)
    step = TrainStep(wl, dev, seed=11, train_mode=False, process_group=dist.group.WORLD, dp_grad_dtype=grad_dtype)
    for _ in range(2):  # second pass exercises the static bucket plan (unused params dropped after step 1)
        step.zero_grad()
        terms = step.forward_backward(feats[sel].contiguous(), labs[sel].contiguous(), Bl)
    torch.cuda.synchronize()
    ce = terms["ce"].detach().clone()
    dist.all_reduce(ce)
    named = list(step.encoder.named_parameters()) + [("head." + n, p) for n, p in step.head.named_parameters()]
    grads = {n: p.grad.detach().cpu().clone() for n, p in named if p.grad is not None}
    mil = terms["mil"].item()
    del terms
    # the same data-parallel step replayed as a CUDA graph (score all-gather + bucketed all-reduces captured)
    from lstc_vad_b200.harness import GraphedTrainStep
    gstep = GraphedTrainStep(step, feats[sel].contiguous(), labs[sel].contiguous(), Bl, warmup=2)
    gterms = gstep()
    gterms = gstep()
    torch.cuda.synchronize()
    graph_err = 0.0
    for n, p in named:
        if p.grad is not None:
            r = grads[n].float()
            graph_err = max(graph_err, ((p.grad.detach().cpu().float() - r).norm() / r.norm().clamp_min(1e-12)).item())
    graph_mil = gterms["mil"].item()
    gstep.close()
    if rank == 0:
        torch.save(dict(mil=mil, ce=ce.item(), grads=grads, graph_err=graph_err, graph_mil=graph_mil), out)
    # the graph (and the collectives it captured) is released: the communicator can be torn down cleanly
    del gstep, gterms
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("grad_dtype,tol", [("fp32", 1e-3), ("bf16", 8e-3)])
def test_two_gpu_data_parallel_step_equals_single_gpu(tmp_path, grad_dtype, tol):
    """fp32 buckets: the exact sum of nn.DataParallel's reduce_add_coalesced.  bf16 buckets (the default payload): every
    rank's gradient is rounded to bf16 once before the sum (relative error <= 2^-9 per element)."""
    from lstc_vad_b200.harness import TrainStep
    out = str(tmp_path / "dp.pt")
    port = 29600 + (os.getpid() % 1000) + (7 if grad_dtype == "bf16" else 0)
    mp.spawn(_worker, args=(2, port, out, grad_dtype), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = _workload()
    feats, labs = _global_batch(wl, dev)
    ref = TrainStep(wl, dev, seed=11, train_mode=False)
    ref.zero_grad()
    terms = ref.forward_backward(feats, labs, wl.batch_size)
    assert abs(got["mil"] - terms["mil"].item()) < 1e-5   # same scores bit for bit -> same global MIL loss
    assert abs(got["ce"] - terms["ce"].item()) < 1e-5
    named = list(ref.encoder.named_parameters()) + [("head." + n, p) for n, p in ref.head.named_parameters()]
    checked = 0
    for n, p in named:
        if p.grad is None:
            assert n not in got["grads"], n
            continue
        g, r = got["grads"][n].float(), p.grad.detach().cpu().float()
        err = (g - r).norm() / r.norm().clamp_min(1e-12)
        # identical kernels on the two halves; only the summation order over windows (and split-K atomics) differs
        assert err < tol, (n, err.item())
        checked += 1
    assert checked > 20
    # CUDA-graph replay of the data-parallel step: same kernels, same collectives
    assert abs(got["graph_mil"] - got["mil"]) < 1e-6
    assert got["graph_err"] < 1e-3, got["graph_err"]


def _opt_worker(rank, world, port, out):
    """Data-parallel STEPS with the fused optimizer on the STN workload shape (odd d_inner 3027: parameter sizes that
    are not multiples of 4 floats inside the flat buckets)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from lstc_vad_b200.harness import TrainStep, Workload, synthetic_step_inputs
    wl = Workload("dp_stn", 256, 3027, 2, 16, 2, 4, relative_pe=False, MHA_layerNorm=False, kind="stn", n_layers=1,
                  n_head=2, d_k=64, dropouts=(0, 0, 0, 0))
    feats, _ = synthetic_step_inputs(wl, seed=4, batch_size=wl.batch_size, device=dev)
    B, P, T = wl.batch_size, wl.part_num, wl.part_len
    Bl = B // world
    nh = B * P * T
    sel = torch.cat([torch.arange(rank * Bl * P * T, (rank + 1) * Bl * P * T),
                     nh + torch.arange(rank * Bl * P * T, (rank + 1) * Bl * P * T)]).to(dev)
    step = TrainStep(wl, dev, seed=13, train_mode=False, process_group=dist.group.WORLD, optimizer=True,
                     dp_grad_dtype="fp32")
    for _ in range(3):
        step.zero_grad()
        step.forward_backward(feats[sel].contiguous(), None, Bl)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({n: p.detach().cpu() for n, p in step.encoder.named_parameters()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_steps_with_fused_adagrad_and_odd_widths(tmp_path):
    from lstc_vad_b200.harness import TrainStep, Workload, synthetic_step_inputs
    out = str(tmp_path / "dp_opt.pt")
    port = 29800 + (os.getpid() % 1000)
    mp.spawn(_opt_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = Workload("dp_stn", 256, 3027, 2, 16, 2, 4, relative_pe=False, MHA_layerNorm=False, kind="stn", n_layers=1,
                  n_head=2, d_k=64, dropouts=(0, 0, 0, 0))
    feats, _ = synthetic_step_inputs(wl, seed=4, batch_size=wl.batch_size, device=dev)
    ref = TrainStep(wl, dev, seed=13, train_mode=False, optimizer=True)
    for _ in range(3):
        ref.zero_grad()
        ref.forward_backward(feats, None, wl.batch_size)
    torch.cuda.synchronize()
    for n, p in ref.encoder.named_parameters():
        r, g = p.detach().cpu().float(), got[n].float()
        assert ((g - r).norm() / r.norm().clamp_min(1e-12)).item() < 1e-4, n
