"""CPU oracle for the LSTC_VAD hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch-CPU restatement of the reference's algorithm, written as pure functions over a
``state_dict`` (the reference is an nn.Module graph).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product path
(``lstc_vad_b200``) never does and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so the oracle is pinned against
the reference itself — ``oracle/make_golden.py`` imports ``/root/reference/models`` and the loss functions
from ``/root/reference/Train/*.py`` in the build container, runs them on seeded inputs and commits the
inputs/outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors (bit-level for integer outputs, 1e-5 for fp32).

Each function cites the reference lines it restates (paths relative to the reference repo root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6  # models/Encoder.py:31, models/MultiHeadAttention.py:47, models/FFN.py:10

# --------------------------------------------------------------------------------------------------
# bf16-faithful mode: the same algorithm, rounded to bfloat16 at exactly the points where the CUDA path stores a
# bf16 tensor (weights, the token matrix after the CLS prepend, q|k|v, post-dropout probabilities, attention output,
# the two residual sums, LayerNorm outputs except the encoder's last, the FFN hidden, the head's first layer).  The
# rounding is a straight-through estimator for autograd, so the gradients are those of the fp32 algorithm evaluated on
# the bf16-rounded forward: the comparison target for a <= 3e-2 bound on the CUDA gradients (the plain fp32 oracle
# stays the stated floor).  Off by default; golden-vector tests always run the fp32 mode.
# --------------------------------------------------------------------------------------------------
_BF16_MODE = False


class bf16_rounding:
    """``with bf16_rounding(): ...`` switches the oracle to the bf16-faithful mode."""

    def __enter__(self):
        global _BF16_MODE
        self.prev, _BF16_MODE = _BF16_MODE, True
        return self

    def __exit__(self, *a):
        global _BF16_MODE
        _BF16_MODE = self.prev


def _round_bf16(g: Tensor) -> Tensor:
    return g.to(torch.bfloat16).to(g.dtype)


def _r(x: Tensor, grad: bool = True) -> Tensor:
    """bf16-faithful mode: x rounded to bf16 (straight-through).  `grad`: the gradient arriving at this tensor in the
    backward is rounded to bf16 as well - wherever the CUDA path keeps an ACTIVATION in bf16 it also hands the
    gradient with respect to it to the next kernel in bf16 (dX, dqkv, dO, the LayerNorm input gradients, dH).  Weights
    and the attention probabilities are forward-only: their gradients stay fp32 (dW, and dP in TMEM)."""
    if not _BF16_MODE:
        return x
    y = x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())
    if grad and y.requires_grad:
        y.register_hook(_round_bf16)
    return y


@dataclass
class EncoderConfig:
    """Mirror of the Encoder constructor arguments that change the math (models/Encoder.py:6-11)."""
    n_layers: int
    n_head: int
    d_k: int
    d_v: int
    d_model: int
    d_inner: int
    MHA_layerNorm: bool = False
    FFN_layerNorm: bool = True
    CLS_learned: bool = False
    position_encoding: bool = False
    relative_pe: bool = False
    window_size: int = 4
    window_depth: int = 3
    input_layerNorm: bool = False
    relative_pe_2D: bool = False
    FFN_need: bool = True


@dataclass
class DropoutMasks:
    """Optional explicit keep-masks (1 = keep) so a run with dropout can be replayed exactly.
    attn[i]: [W,H,L,L] ; fc[i]: [W,L,D] ; ffn[i]: [W,L,D] ; pos: [W,L,D]; *_p are the drop probabilities."""
    attn: Dict[int, Tensor] = field(default_factory=dict)
    fc: Dict[int, Tensor] = field(default_factory=dict)
    ffn: Dict[int, Tensor] = field(default_factory=dict)
    pos: Optional[Tensor] = None
    attn_p: float = 0.0
    fc_p: float = 0.0
    ffn_p: float = 0.0
    pos_p: float = 0.0


def _apply_mask(x: Tensor, keep: Optional[Tensor], p: float) -> Tensor:
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype).reshape(x.shape) / (1.0 - p)


# --------------------------------------------------------------------------------------------------
# relative-position index buffers (integer work: bit-exact)
# --------------------------------------------------------------------------------------------------
def relative_position_index_3d(window_depth: int, window_size: int) -> np.ndarray:
    """int64 [(wd*ws*ws), (wd*ws*ws)] — models/MultiHeadAttention.py:59-73.
    Token order is depth-major, then row, then column; the index of pair (a, b) is
    ((da-db)+wd-1)*(2ws-1)^2 + ((ha-hb)+ws-1)*(2ws-1) + ((wa-wb)+ws-1)."""
    wd, ws = window_depth, window_size
    d, h, w = np.meshgrid(np.arange(wd), np.arange(ws), np.arange(ws), indexing="ij")
    d, h, w = d.reshape(-1), h.reshape(-1), w.reshape(-1)
    span = 2 * ws - 1
    rel = ((d[:, None] - d[None, :]) + wd - 1) * span * span
    rel = rel + ((h[:, None] - h[None, :]) + ws - 1) * span
    rel = rel + ((w[:, None] - w[None, :]) + ws - 1)
    return rel.astype(np.int64)


def relative_position_index_2d(window_size: int) -> np.ndarray:
    """int64 [ws*ws, ws*ws] — models/MultiHeadAttention.py:76-89."""
    ws = window_size
    h, w = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    h, w = h.reshape(-1), w.reshape(-1)
    span = 2 * ws - 1
    rel = ((h[:, None] - h[None, :]) + ws - 1) * span + ((w[:, None] - w[None, :]) + ws - 1)
    return rel.astype(np.int64)


def relative_bias(table: Tensor, index: Tensor, n_tokens: int) -> Tensor:
    """table [T,H], index [n,n] -> bias [H, n_tokens, n_tokens] for the non-CLS block
    (models/MultiHeadAttention.py:108-110: the index is sliced to [:len_q-1, :len_q-1])."""
    idx = index[:n_tokens, :n_tokens].reshape(-1)
    return table[idx].reshape(n_tokens, n_tokens, -1).permute(2, 0, 1)


# --------------------------------------------------------------------------------------------------
# model forward
# --------------------------------------------------------------------------------------------------
def cls_prepend(x: Tensor, cls_token: Optional[Tensor] = None) -> Tensor:
    """[W,L0,D] -> [W,L0+1,D]: CLS = token mean, or the learned token (models/Encoder.py:51-55,
    models/PatchEmbedding.py:12-19)."""
    if cls_token is not None:
        cls = cls_token.reshape(1, 1, -1).expand(x.shape[0], 1, x.shape[2])
    else:
        cls = x.mean(dim=1, keepdim=True)
    return torch.cat([cls, x], dim=1)


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * weight + bias


def mha_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, cfg: EncoderConfig, layer: int = 0,
                masks: Optional[DropoutMasks] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """models/MultiHeadAttention.py:93-126 for self-attention (q = k = v = x, mask = None).
    Returns (out [W,L,D], attn [W,H,L,L] post-dropout, v [W,H,L,dv])."""
    W, L, _ = x.shape
    H, dk, dv = cfg.n_head, cfg.d_k, cfg.d_v
    q = _r(x @ _r(sd[prefix + "w_qs.weight"], False).t()).reshape(W, L, H, dk).permute(0, 2, 1, 3)
    k = _r(x @ _r(sd[prefix + "w_ks.weight"], False).t()).reshape(W, L, H, dk).permute(0, 2, 1, 3)
    v = _r(x @ _r(sd[prefix + "w_vs.weight"], False).t()).reshape(W, L, H, dv).permute(0, 2, 1, 3)
    s = (q / (dk ** 0.5)) @ k.transpose(2, 3)  # :103 — scale applied to q before the product
    if cfg.relative_pe:  # :107-111
        b = relative_bias(sd[prefix + "relative_position_bias_table"], sd[prefix + "relative_position_index"], L - 1)
        pad = torch.zeros(H, L, L, dtype=s.dtype)
        pad = torch.cat([torch.zeros(H, 1, L, dtype=s.dtype),
                         torch.cat([torch.zeros(H, L - 1, 1, dtype=s.dtype), b], dim=2)], dim=1)
        s = s + pad.unsqueeze(0)
    if cfg.relative_pe_2D:  # :113-117 — uses the full index: requires L-1 == ws*ws
        n = cfg.window_size * cfg.window_size
        b = relative_bias(sd[prefix + "relative_position_bias_table"], sd[prefix + "relative_position_index"], n)
        pad = torch.cat([torch.zeros(H, 1, L, dtype=s.dtype),
                         torch.cat([torch.zeros(H, L - 1, 1, dtype=s.dtype), b], dim=2)], dim=1)
        s = s + pad.unsqueeze(0)
    if _BF16_MODE and s.requires_grad:
        s.register_hook(_round_bf16)  # the CUDA backward stores (scale * dS) as a bf16 operand of dQ / dK
    attn = torch.softmax(s, dim=-1)
    if masks is not None:
        attn = _apply_mask(attn, masks.attn.get(layer), masks.attn_p)  # :119
    o = _r(_r(attn, False) @ v).permute(0, 2, 1, 3).reshape(W, L, H * dv)  # :120-122
    o = o @ _r(sd[prefix + "fc.weight"], False).t()
    if masks is not None:
        o = _apply_mask(o, masks.fc.get(layer), masks.fc_p)  # :123
    o = _r(o + x)  # :124
    if cfg.MHA_layerNorm:
        o = _r(layer_norm(o, sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"]))  # :125-126
    return o, attn, v


def ffn_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, cfg: EncoderConfig, layer: int = 0,
                masks: Optional[DropoutMasks] = None, last: bool = False) -> Tensor:
    """models/FFN.py:14-22.  `last`: the encoder's final block (its LayerNorm output stays fp32 in the CUDA path)."""
    h = _r(torch.relu(x @ _r(sd[prefix + "w_1.weight"], False).t() + sd[prefix + "w_1.bias"]))
    y = h @ _r(sd[prefix + "w_2.weight"], False).t() + sd[prefix + "w_2.bias"]
    if masks is not None:
        y = _apply_mask(y, masks.ffn.get(layer), masks.ffn_p)
    y = _r(y + x)
    if cfg.FFN_layerNorm:
        y = layer_norm(y, sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])
        if not last:
            y = _r(y)
    return y


def encoder_forward(sd: Dict[str, Tensor], x: Tensor, cfg: EncoderConfig, masks: Optional[DropoutMasks] = None,
                    return_all: bool = False):
    """models/Encoder.py:43-74 (+ models/EncoderLayer.py:18-30).  x [W,L0,D] -> [W,L0+1,D]
    (and the per-layer attention / value lists when return_all)."""
    if cfg.input_layerNorm:
        x = layer_norm(x, sd["layer_norm.weight"], sd["layer_norm.bias"])
    x = cls_prepend(x, sd["cls_token"] if cfg.CLS_learned else None)
    if cfg.position_encoding:
        x = x + sd["position_enc"][:, : x.shape[1], :]
        if masks is not None:
            x = _apply_mask(x, masks.pos, masks.pos_p)
    x = _r(x)
    attns: List[Tensor] = []
    vs: List[Tensor] = []
    for i in range(cfg.n_layers):
        x, a, v = mha_forward(sd, f"layer_stack.{i}.slf_attn.", x, cfg, i, masks)
        if cfg.FFN_need:
            x = ffn_forward(sd, f"layer_stack.{i}.pos_ffn.", x, cfg, i, masks, last=(i == cfg.n_layers - 1))
        attns.append(a)
        vs.append(v)
    if return_all:
        return x, attns, vs
    return x


def head_forward(sd: Dict[str, Tensor], x: Tensor, kind: str, masks: Optional[Tuple[Tensor, Tensor]] = None,
                 p: float = 0.0) -> Tensor:
    """Classifier (models/Classifier.py:8-23, kind='classifier', softmax over 2 classes) or Regressor
    (models/Regressor.py:7-20, kind='regressor', sigmoid).  ReLU only after the first Linear."""
    x = _r(x.reshape(-1, x.shape[-1]))
    h = torch.relu(x @ _r(sd[f"{kind}.0.weight"], False).t() + sd[f"{kind}.0.bias"])
    if masks is not None:
        h = _apply_mask(h, masks[0], p)
    h = _r(h)
    h = h @ sd[f"{kind}.3.weight"].t() + sd[f"{kind}.3.bias"]
    if masks is not None:
        h = _apply_mask(h, masks[1], p)
    z = h @ sd[f"{kind}.5.weight"].t() + sd[f"{kind}.5.bias"]
    return torch.softmax(z, dim=-1) if kind == "classifier" else torch.sigmoid(z)


# --------------------------------------------------------------------------------------------------
# losses and labelling
# --------------------------------------------------------------------------------------------------
def mil_loss(y_pred: Tensor, batch_size: int, part_num: int, part_len: int = 1, lambda_1: float = 0.01,
             topk: int = 1) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """get_MIL_loss — LTN form Train/temporal_transformer_shanghaitech.py:25-36 (part_len=1, flat y_pred) and
    STN form Train/spatio_transformer_shanghaitech.py:21-32 (y_pred [2B, P*T, 1]).
    The sparsity term is ``mean(y_pred[B:])`` on y_pred AS SHAPED BY THE CALLER: for a flat vector that is
    "all but the first B scores", for [2B, ...] it is the abnormal half.
    Returns (loss, err, spar, argmax index per bag)."""
    B = batch_size
    part = y_pred.reshape(2 * B, part_num, part_len).mean(-1)
    if topk == 1:
        bag, idx = part.max(dim=-1)
    else:
        vals, idx = part.topk(topk, dim=-1)
        bag = vals.mean(-1)
    nor, abn = bag[:B], bag[B:]
    err = torch.relu(1.0 - abn[None, :] + nor[:, None]).sum() / (B * B)
    spar = y_pred[B:].mean()
    return err + lambda_1 * spar, err, spar, idx


def ce_loss(outputs: Tensor, labels: Tensor) -> Tensor:
    """get_CE_loss — Train/temporal_transformer_shanghaitech.py:21-23: F.cross_entropy with probability
    targets applied to the classifier's softmax OUTPUT (so a second log-softmax is taken)."""
    return -(labels * torch.log_softmax(outputs, dim=-1)).sum(-1).mean()


def bce_loss(outputs: Tensor, labels: Tensor, lambda_normal: float, lambda_abnormal: float) -> Tensor:
    """get_BCE_loss — Train/spatio_transformer_MIL_CE.py:23-26.  outputs [2B,P], labels [2B,P,2]."""
    return torch.mean(-lambda_normal * labels[..., 0] * torch.log(1 - outputs + 1e-8)
                      - lambda_abnormal * labels[..., 1] * torch.log(outputs + 1e-8))


def soft_labels(abnormal_clip_labels: Tensor, batch_size: int, part_num: int, part_len: int) -> Tensor:
    """Train/temporal_transformer_shanghaitech.py:103-112: normal windows -> (1,0); abnormal windows ->
    (1-m, m) with m the mean pseudo label over the part_len clips.  Returns [2*B*P, 2], normal first."""
    m = abnormal_clip_labels.reshape(batch_size, part_num, part_len).float().mean(-1)
    abn = torch.stack([1.0 - m, m], dim=-1)
    nor = torch.zeros(batch_size, part_num, 2)
    nor[..., 0] = 1.0
    return torch.cat([nor, abn], dim=0).reshape(2 * batch_size * part_num, 2)


def threshold_labels(scores: Tensor, thr: float) -> Tensor:
    """Train/pseudo_labels_generator_temporal.py:103-104,139-140: where(s > thr, s, 0) (strict >)."""
    return torch.where(scores > thr, scores, torch.zeros_like(scores))


def ltn_train_loss(enc_sd, cls_sd, feats: Tensor, clip_labels: Tensor, cfg: EncoderConfig, batch_size: int,
                   part_num: int, lambda_1=0.01, lambda_MIL=1.0, lambda_CE=0.8):
    """The loss of one LTN train step, Train/temporal_transformer_shanghaitech.py:122-134.
    feats [2BP, T*N, D] (normal windows first); clip_labels [2BP, 2]."""
    out = encoder_forward(enc_sd, feats, cfg)
    probs = head_forward(cls_sd, out[:, 0, :], "classifier")
    ce = ce_loss(probs, clip_labels)
    mil, err, spar, idx = mil_loss(probs[:, 1], batch_size, part_num, 1, lambda_1)
    return lambda_MIL * mil + lambda_CE * ce, dict(mil=mil, err=err, spar=spar, ce=ce, probs=probs, idx=idx)


def stn_train_loss(enc_sd, reg_sd, feats: Tensor, cfg: EncoderConfig, batch_size: int, part_num: int, part_len: int,
                   lambda_1=0.01):
    """One STN train step's loss, Train/spatio_transformer_shanghaitech.py:90-101.
    feats [2*B*P*T, N, D]; the regressor output is viewed [2B, P*T, 1] before get_MIL_loss."""
    out = encoder_forward(enc_sd, feats, cfg)
    score = head_forward(reg_sd, out[:, 0, :], "regressor").reshape(2 * batch_size, part_num * part_len, 1)
    mil, err, spar, idx = mil_loss(score, batch_size, part_num, part_len, lambda_1)
    return mil, dict(err=err, spar=spar, scores=score, idx=idx)


# --------------------------------------------------------------------------------------------------
# window policies of the eval / labelling loops (used by the sharded-inference tests)
# --------------------------------------------------------------------------------------------------
def window_bounds(n_clips: int, part_len: int, backshift: bool) -> List[Tuple[int, int]]:
    """Window [beg, end) list for one video.  backshift=False: the last window is short
    (Train/pseudo_labels_generator_temporal.py:127-134); backshift=True: the last window is moved back to
    full length (Test/evaluation_shanghaitech_ubnormal.py:74-86)."""
    out = []
    n_win = int(math.ceil(n_clips / part_len))
    for i in range(n_win):
        beg, end = i * part_len, min((i + 1) * part_len, n_clips)
        if backshift and end - beg < part_len and n_clips >= part_len:
            out.append((end - part_len, end))
        else:
            out.append((beg, end))
    return out


def pool_video_bins(feats: Tensor, n_bins: int = 32, l2norm: bool = True):
    """Test/evaluation_UCF.py:52-77: r = np.linspace(0, n_clips, n_bins + 1, dtype=int32); bin b is the mean of clips
    [r[b], r[b+1]) — or the single clip r[b] when the bin is empty — then F.normalize(p=2, dim=-1) per token.
    feats [n_clips, n_patch, D] -> ([n_bins, n_patch, D], r)."""
    n_clips = feats.shape[0]
    r = np.linspace(0, n_clips, n_bins + 1, dtype=np.int32)
    rows = []
    for b in range(n_bins):
        if r[b] == r[b + 1]:
            rows.append(feats[r[b]])
        else:
            rows.append(feats[r[b]:r[b + 1]].mean(dim=0))
    out = torch.stack(rows)
    if l2norm:
        out = F.normalize(out, p=2, dim=-1)
    return out, r.tolist()


def sample_window_indices(feat_len: int, part_num: int, part_len: int, sample: str, rng):
    """utils/load_dataset.py:68-88 (`sample_feat`): the clip indices `chosen[:part_num*part_len]`.
    'uniform': linspace(0, len-T, P+1) shifted by ONE random offset in [0, (len-T)//(P+1)); otherwise every window gets
    its own random offset in [0, spacing).  `rng` is np.random or a RandomState (same call order as the reference)."""
    base = np.linspace(0, feat_len - part_len, num=part_num + 1, dtype=int)
    ar = np.arange(0, part_len, 1, dtype=int)
    if sample == "uniform":
        span = (feat_len - part_len) // (part_num + 1)
        move = 0 if span < 1 else rng.randint(span)
        chosen = (base + move).repeat(part_len).reshape([-1, part_len]) + ar
    else:
        chosen = base.repeat(part_len).reshape([-1, part_len]) + ar
        gap = chosen[1, 0] - chosen[0, 0]
        move = 0 if gap == 0 else rng.randint(0, gap, [part_num + 1]).repeat(part_len).reshape([-1, part_len])
        chosen = chosen + move
    return chosen.reshape([-1])[: part_num * part_len]
