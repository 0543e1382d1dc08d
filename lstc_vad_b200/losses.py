"""Loss functions of the reference train scripts, same names / signatures / return values, on CUDA kernels.

The reference defines these inside each `Train/*.py` script; `args` is the script's argparse namespace
(only the attributes read by the reference are read here):
  get_MIL_loss(args, y_pred[, part_len])  Train/temporal_transformer_shanghaitech.py:25-36 (LTN form),
                                          Train/spatio_transformer_shanghaitech.py:21-32 and
                                          Train/spatio_transformer_MIL_CE.py:32-44 (STN forms)
  get_CE_loss(args, outputs, labs)        Train/temporal_transformer_shanghaitech.py:21-23
  get_BCE_loss(args, outputs, labs)       Train/spatio_transformer_MIL_CE.py:23-26
  threshold_pseudo_labels(scores, thr)    Train/pseudo_labels_generator_temporal.py:103-104
  soft_clip_labels(...)                   Train/temporal_transformer_shanghaitech.py:103-112
"""
from __future__ import annotations

import torch

from . import functional as Fn
from . import ops


def get_MIL_loss(args, y_pred, part_len=None):
    """MIL ranking + sparsity loss; returns (loss, err, spar_l1) like the reference.

    * LTN form (`part_len` omitted and y_pred flat [2*B*P]): bag score = max over the P window scores.
    * STN form (y_pred [2B, P*T, 1], T = `part_len` argument or args.part_len): window score = mean over T.
    The sparsity term is mean(y_pred[B:]) on y_pred *as shaped by the caller* — for a flat vector that is every
    score but the first B (a quirk of the reference kept on purpose).  k of the top-k is 1 (`torch.max`); ties
    resolve to the lowest index.  Only `loss` carries a gradient (err / spar_l1 are logging values)."""
    B, P = int(args.batch_size), int(args.part_num)
    n = y_pred.numel()
    if part_len is None:
        T = n // (2 * B * P)
        if T != 1 and hasattr(args, "part_len"):
            T = int(args.part_len)
    else:
        T = int(part_len)
    if 2 * B * P * T != n:
        raise RuntimeError(f"get_MIL_loss: y_pred has {n} scores, expected 2*B*P*T = {2 * B * P * T}")
    rows_per_slice = n // y_pred.shape[0] if y_pred.dim() > 0 else 1
    spar_start = min(B, y_pred.shape[0]) * rows_per_slice  # y_pred[B:] along dim 0
    topk = 1
    loss, err, spar, _ = Fn.MILLossFn.apply(y_pred, B, P, T, topk, float(getattr(args, "lambda_1", 0.01)), spar_start)
    return loss, err, spar


def mil_topk_indices(args, y_pred, part_len=1, topk=1):
    """The per-bag selected window indices (int32 [2B, topk]) — bit-exact whenever the scores are."""
    B, P = int(args.batch_size), int(args.part_num)
    _, idx, _ = ops.mil_loss(y_pred.detach().reshape(-1).float().contiguous(), B, P, int(part_len), topk,
                             float(getattr(args, "lambda_1", 0.01)), B, need_grad=False)
    return idx


def get_CE_loss(args, outputs, labs):
    """F.cross_entropy(outputs, labs) with probability targets, applied — as the reference does — to the
    classifier's softmax OUTPUT."""
    return Fn.SoftCEFn.apply(outputs, labs)


def get_BCE_loss(args, outputs, labs):
    """Weighted BCE on per-part mean regressor scores: outputs [2B,P], labs [2B,P,2]."""
    return Fn.BCEFn.apply(outputs, labs, 1, float(args.lambda_normal), float(args.lambda_abnormal))


def threshold_pseudo_labels(scores: torch.Tensor, threshold: float) -> torch.Tensor:
    """where(score > threshold, score, 0) — keeps the soft value, strict '>'."""
    return ops.threshold_labels(scores.detach().float(), threshold).view(scores.shape)


def soft_clip_labels(abnormal_clip_labels: torch.Tensor, batch_size: int, part_num: int, part_len: int) -> torch.Tensor:
    """[B, P*T(,1)] pseudo labels of the abnormal videos -> [2*B*P, 2] soft targets, normal windows first.
    Host-side glue (a few KB); plain torch on whatever device the labels live on."""
    m = abnormal_clip_labels.reshape(batch_size, part_num, part_len).float().mean(-1)
    abn = torch.stack([1.0 - m, m], dim=-1)
    nor = torch.zeros(batch_size, part_num, 2, device=m.device)
    nor[..., 0] = 1.0
    return torch.cat([nor, abn], dim=0).reshape(2 * batch_size * part_num, 2)
