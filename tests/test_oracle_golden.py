"""Pins the CPU oracle against golden vectors produced by the unmodified reference (oracle/make_golden.py).

CPU-only.  fp32 tolerances are 1e-5 relative (the oracle reorders a few fp32 sums, e.g. the hinge loop);
integer outputs (relative-position indices, argmax indices, thresholded labels) must be bit-exact.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import lstc_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return torch.load(GOLD / f"{name}.pt", weights_only=False)


def cfg_from(kw):
    fields = O.EncoderConfig.__dataclass_fields__.keys()
    return O.EncoderConfig(**{k: v for k, v in kw.items() if k in fields})


def close(a, b, tol=1e-5):
    a, b = a.detach().double(), b.detach().double()
    err = (a - b).abs().max().item()
    scale = max(1.0, b.abs().max().item())
    assert err <= tol * scale, f"max abs err {err:.3e} (scale {scale:.3e})"


def test_relative_position_indices_bit_exact():
    idx = load("relpos_index")
    for ws, wd in ((4, 3), (4, 2), (4, 5), (3, 3)):
        mine = torch.from_numpy(O.relative_position_index_3d(wd, ws))
        assert mine.dtype == torch.int64
        assert torch.equal(mine, idx[f"3d_ws{ws}_wd{wd}"])
    for ws in (4, 3):
        assert torch.equal(torch.from_numpy(O.relative_position_index_2d(ws)), idx[f"2d_ws{ws}"])


@pytest.mark.parametrize("name", ["ltn_relpe", "ltn_ucf_sliced"])
def test_ltn_forward_loss_and_grads(name):
    c = load(name)
    cfg = cfg_from(c["enc_kwargs"])
    enc_sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["enc_state"].items()}
    cls_sd = {k: v.clone().requires_grad_(True) for k, v in c["cls_state"].items()}
    x = c["x"].clone().requires_grad_(True)
    out, attns, vs = O.encoder_forward(enc_sd, x, cfg, return_all=True)
    close(out, c["enc_out"])
    close(attns[0], c["attn0"])
    close(vs[0], c["v0"])
    labs = O.soft_labels(c["clip_pseudo"], c["B"], c["P"], c["T"])
    close(labs, c["clip_labs"], 1e-7)
    loss, aux = O.ltn_train_loss(enc_sd, cls_sd, x, labs, cfg, c["B"], c["P"])
    close(aux["probs"], c["probs"])
    for k in ("ce", "mil", "err", "spar"):
        close(aux[k], c[k])
    close(loss, c["loss"])
    assert torch.equal(aux["idx"], c["topk_idx"])
    assert torch.equal(O.threshold_labels(aux["probs"][:, 1].detach(), 0.5), c["thr_labels"]) or \
        (O.threshold_labels(c["probs"][:, 1], 0.5) == c["thr_labels"]).all()
    loss.backward()
    close(x.grad, c["x_grad"], 1e-4)
    for k, g in c["enc_grads"].items():
        if g is None:
            assert enc_sd[k].grad is None or enc_sd[k].grad.abs().max() == 0, k
        else:
            close(enc_sd[k].grad, g, 1e-4)
    for k, g in c["cls_grads"].items():
        close(cls_sd[k].grad, g, 1e-4)
    with torch.no_grad():
        for L0, v in c["var_L"].items():
            o = O.encoder_forward(enc_sd, v["x"], cfg)
            close(o, v["enc_out"])
            close(O.head_forward(cls_sd, o[:, 0, :], "classifier"), v["probs"])


@pytest.mark.parametrize("name", ["stn_plain", "stn_relpe2d_cls_pos"])
def test_stn_forward_loss_and_grads(name):
    c = load(name)
    cfg = cfg_from(c["enc_kwargs"])
    enc_sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["enc_state"].items()}
    reg_sd = {k: v.clone().requires_grad_(True) for k, v in c["reg_state"].items()}
    x = c["x"].clone().requires_grad_(True)
    B, P, T = c["B"], c["P"], c["T"]
    mil, aux = O.stn_train_loss(enc_sd, reg_sd, x, cfg, B, P, T)
    close(aux["scores"], c["scores"])
    close(mil, c["mil"])
    close(aux["err"], c["err"])
    close(aux["spar"], c["spar"])
    part = aux["scores"].reshape(2 * B, P, T).mean(-1)
    bce = O.bce_loss(part, c["bce_labs"], 0.2, 2.0)
    close(bce, c["bce"])
    (mil + 0.5 * bce).backward()
    close(x.grad, c["x_grad"], 1e-4)
    for k, g in c["enc_grads"].items():
        if g is None:
            assert enc_sd[k].grad is None or enc_sd[k].grad.abs().max() == 0, k
        else:
            close(enc_sd[k].grad, g, 1e-4)
    for k, g in c["reg_grads"].items():
        close(reg_sd[k].grad, g, 1e-4)


def test_mil_sparsity_slice_quirk():
    """LTN form: a flat score vector is sliced [B:], i.e. all but the first B WINDOWS."""
    B, P = 4, 8
    y = torch.arange(2 * B * P, dtype=torch.float32) / (2 * B * P)
    _, _, spar, _ = O.mil_loss(y, B, P, 1)
    assert torch.isclose(spar, y[B:].mean())
    _, _, spar2, _ = O.mil_loss(y.reshape(2 * B, P, 1), B, P, 1)
    assert torch.isclose(spar2, y[B * P:].mean())


def test_window_bounds_policies():
    assert O.window_bounds(7, 3, backshift=False) == [(0, 3), (3, 6), (6, 7)]
    assert O.window_bounds(7, 3, backshift=True) == [(0, 3), (3, 6), (4, 7)]
    assert O.window_bounds(6, 3, backshift=True) == [(0, 3), (3, 6)]
    assert O.window_bounds(2, 3, backshift=True) == [(0, 2)]


def test_window_sampler_indices_match_reference_sample_feat():
    """Golden: clip indices chosen by the reference's own `sample_feat` (utils/load_dataset.py:56-88) under a seeded
    np.random; both the oracle restatement and the product's host function must reproduce them bit for bit."""
    from lstc_vad_b200.harness import sample_window_indices
    cases = torch.load(GOLD / "window_sampler.pt", weights_only=False)
    assert len(cases) == 24
    for c in cases:
        np.random.seed(c["seed"])
        mine = O.sample_window_indices(c["n"], c["P"], c["T"], c["sample"], np.random)
        assert np.array_equal(mine, c["chosen"]), c
        np.random.seed(c["seed"])
        prod = sample_window_indices(c["n"], c["P"], c["T"], c["sample"], np.random)
        assert np.array_equal(prod, c["chosen"]), c
