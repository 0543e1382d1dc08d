"""CPU, world_size 2, gloo: the data-parallel sharding scheme of harness.TrainStep is exact.

Bags are sharded by video pair; every rank all-gathers the per-window scores, evaluates the GLOBAL MIL loss and
keeps the gradient slice of its own windows; CE is a share of the global mean; parameter gradients are SUM
all-reduced.  Checked here with the CPU oracle standing in for the kernels: the sharded loss and the reduced
parameter gradients must equal the single-process result on the whole batch (incl. the flat-slice sparsity quirk,
whose excluded "first B windows" all live on rank 0)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lstc_oracle as O

WORLD = 2
KW = dict(n_layers=1, n_head=2, d_k=16, d_v=16, d_model=32, d_inner=64, MHA_layerNorm=True, FFN_layerNorm=True,
          relative_pe=True, window_size=2, window_depth=2)
B, P, T, N, D = 4, 3, 2, 4, 32  # global batch: 4 video pairs


def make_problem():
    import sys
    sys.path.insert(0, "/root/reference") if False else None
    from lstc_vad_b200.models import Classifier, Encoder
    torch.manual_seed(0)
    enc = Encoder(weight_init=False, **KW)
    cls = Classifier(D, 0.6)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2 * B * P, T * N, D, generator=g).abs()
    labs = O.soft_labels(torch.rand(B, P * T, generator=g), B, P, T)
    return enc.state_dict(), cls.state_dict(), x, labs


def as_leaf(sd):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}


def worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from lstc_vad_b200.harness import global_score_order, local_grad_slice
    esd0, csd0, x, labs = make_problem()
    cfg = O.EncoderConfig(**{k: v for k, v in KW.items() if k in O.EncoderConfig.__dataclass_fields__})
    Bl = B // WORLD
    nh = B * P  # windows in the global normal half
    sel = torch.cat([torch.arange(rank * Bl * P, (rank + 1) * Bl * P), nh + torch.arange(rank * Bl * P, (rank + 1) * Bl * P)])
    esd, csd = as_leaf(esd0), as_leaf(csd0)
    out_l = O.encoder_forward(esd, x[sel], cfg)
    probs = O.head_forward(csd, out_l[:, 0, :], "classifier")
    score = probs[:, 1].contiguous()
    gathered = [torch.empty_like(score) for _ in range(WORLD)]
    dist.all_gather(gathered, score.detach())
    glob = global_score_order(gathered).requires_grad_(True)
    mil, err, spar, _ = O.mil_loss(glob, B, P, 1, 0.01)
    mil.backward()
    d_local = local_grad_slice(glob.grad, rank, WORLD)
    ce_share = O.ce_loss(probs, labs[sel]) / WORLD
    (score * d_local).sum().backward(retain_graph=True)   # inject the MIL gradient slice
    (0.8 * ce_share).backward()
    ce = ce_share.detach().clone()
    dist.all_reduce(ce)
    grads = {}
    for k, v in list(esd.items()) + [("cls." + k, v) for k, v in csd.items()]:
        if v.requires_grad and v.grad is not None:
            gsum = v.grad.clone()
            dist.all_reduce(gsum)
            grads[k] = gsum
    if rank == 0:
        torch.save(dict(loss=(mil.detach() + 0.8 * ce), grads=grads), out)
    dist.destroy_process_group()


def test_sharded_loss_and_gradients_equal_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(port, out), nprocs=WORLD, join=True)
    got = torch.load(out, weights_only=False)
    esd0, csd0, x, labs = make_problem()
    cfg = O.EncoderConfig(**{k: v for k, v in KW.items() if k in O.EncoderConfig.__dataclass_fields__})
    esd, csd = as_leaf(esd0), as_leaf(csd0)
    loss, _ = O.ltn_train_loss(esd, csd, x, labs, cfg, B, P)
    loss.backward()
    assert torch.allclose(got["loss"], loss.detach(), atol=1e-6)
    for k, v in list(esd.items()) + [("cls." + k, v) for k, v in csd.items()]:
        if v.requires_grad and v.grad is not None:
            assert torch.allclose(got["grads"][k], v.grad, atol=1e-5, rtol=1e-4), k


def test_score_order_round_trip():
    from lstc_vad_b200.harness import global_score_order, local_grad_slice
    per_rank = [torch.arange(6, dtype=torch.float32) + 100 * r for r in range(4)]
    glob = global_score_order(per_rank)
    assert glob[:12].tolist() == [0, 1, 2, 100, 101, 102, 200, 201, 202, 300, 301, 302]
    for r in range(4):
        assert torch.equal(local_grad_slice(glob, r, 4), per_rank[r])
