"""Autograd glue: one torch.autograd.Function per fused block of the reference graph, each with a hand-written
backward made of the same C-ABI kernels (no torch math on the hot path).

Blocks and the reference code they replace:
  ClsPrependFn   models/Encoder.py:51-59            CLS token (+ abs. position encoding, dropout)
  LayerNormFn    models/Encoder.py:48-49            stand-alone input LayerNorm
  MHABlockFn     models/MultiHeadAttention.py:93-126 QKV proj, attention, out-proj, dropout, residual, LN
  FFNBlockFn     models/FFN.py:14-22                w_1, ReLU, w_2, dropout, residual, LN
  HeadFn         models/Classifier.py:8-23, models/Regressor.py:7-20
  MILLossFn / SoftCEFn / BCEFn                      Train/*.py loss functions (see losses.py)

Activations between blocks are bf16, parameters stay fp32 in the state_dict (bf16 copies are cached per
parameter version), GEMMs accumulate in fp32, LayerNorm / softmax / losses compute in fp32.
"""
from __future__ import annotations

import os

import threading
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .ops import BF16, F32, NO_DROPOUT

# --------------------------------------------------------------------------------------------------
# dropout stream: (seed, offset) pairs; one fresh offset per dropout site per call
# --------------------------------------------------------------------------------------------------
_rng_lock = threading.Lock()
_rng_state = {"seed": None, "offset": 0}


_rng_counters = {}  # device index -> int64[1] tensor registered with the library (lives for the whole process)


def device_rng_counter(device=None) -> torch.Tensor:
    """The device-resident dropout step counter of `device`, shared by every CUDA graph of the process.

    Every mask-drawing kernel adds its value to the `offset` argument it was launched with, so a replayed graph (which
    bakes seed / offset by value) draws fresh masks as long as the graph bumps the counter.  ONE counter per device is
    allocated on first use, registered with `lstc_set_rng_step` and never freed or re-pointed: the registration is a
    per-device __constant__ pointer inside the library, so per-graph counters would silently re-point each other's
    kernels (and a freed one would dangle)."""
    from . import _lib
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _rng_lock:
        t = _rng_counters.get(idx)
        if t is None:
            with torch.cuda.device(idx):
                t = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", idx))
                _lib.check(_lib.load().lstc_set_rng_step(t.data_ptr()), "lstc_set_rng_step")
            _rng_counters[idx] = t
        return t


def set_dropout_stream(seed: int, offset: int = 0) -> None:
    """Pins the dropout stream (e.g. seed + rank for data-parallel replicas)."""
    with _rng_lock:
        _rng_state["seed"] = int(seed) & 0x7FFFFFFFFFFFFFFF
        _rng_state["offset"] = int(offset)


def next_dropout(p: float, training: bool):
    if not training or p <= 0.0:
        return NO_DROPOUT
    if p >= 1.0:
        raise ValueError("dropout probability must be < 1")
    with _rng_lock:
        if _rng_state["seed"] is None:
            _rng_state["seed"] = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        _rng_state["offset"] += 1
        return (float(p), _rng_state["seed"], _rng_state["offset"])


# --------------------------------------------------------------------------------------------------
# bf16 weight cache, keyed on the parameter's identity + version (optimizer steps bump _version)
# --------------------------------------------------------------------------------------------------
class WeightCache:
    """bf16 (optionally zero-padded / concatenated) copies of fp32 parameters.

    Only genuine nn.Parameters are cached: the per-forward broadcast copies that nn.DataParallel puts in
    replicas are plain tensors whose storage address can be recycled with new contents, so they are
    re-cast on every call.  An in-place update through ``param.data`` does not bump ``_version``; call
    ``invalidate()`` after such an update."""

    def __init__(self):
        self._d = {}
        self._lock = threading.Lock()

    def invalidate(self):
        with self._lock:
            self._d.clear()

    @staticmethod
    def _sig(p: torch.Tensor):
        return (p.data_ptr(), p._version, tuple(p.shape), p.device.index)

    def _get(self, key, params, build):
        cacheable = all(isinstance(p, torch.nn.Parameter) for p in params)
        sig = tuple(self._sig(p) for p in params)
        if cacheable:
            with self._lock:
                hit = self._d.get(key)
            # identity is checked through weak references: `id()` and the storage address of a dead parameter can
            # both be recycled by a new one with the same shape and version
            if hit is not None and hit[0] == sig and all(r() is p for r, p in zip(hit[2], params)):
                return hit[1]
        val = build()
        if cacheable:
            with self._lock:
                if len(self._d) > 4096:  # entries of dead modules
                    self._d = {k: v for k, v in self._d.items() if all(r() is not None for r in v[2])}
                self._d[key] = (sig, val, tuple(weakref.ref(p) for p in params))
        return val

    def bf16(self, p: torch.Tensor, pad_rows: int = 0, pad_cols: int = 0) -> torch.Tensor:
        """p fp32 [R,C] -> bf16 [R+pad_rows, C+pad_cols] (zero padding)."""
        def build():
            src = p.detach()
            if pad_rows == 0 and pad_cols == 0:
                return ops.cast_to_bf16(src)
            R, C = src.shape
            buf = torch.zeros((R + pad_rows, C + pad_cols), device=src.device, dtype=BF16)
            if pad_cols == 0:
                ops.cast_to_bf16(src, out=buf[:R])
            else:
                buf[:R, :C].copy_(src)  # odd-width (e.g. d_inner 3027) one-off repack per weight version
            return buf
        return self._get(("bf16", id(p), p.device.index, pad_rows, pad_cols), (p,), build)

    def qkv(self, wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
        """[w_qs; w_ks; w_vs] stacked into one bf16 [3*H*dk, D] matrix (one GEMM for the three projections)."""
        def build():
            n = wq.shape[0]
            buf = torch.empty((3 * n, wq.shape[1]), device=wq.device, dtype=BF16)
            for i, w in enumerate((wq, wk, wv)):
                ops.cast_to_bf16(w.detach(), out=buf[i * n:(i + 1) * n])
            return buf
        return self._get(("qkv", id(wq), id(wk), id(wv), wq.device.index), (wq, wk, wv), build)

    def bf16_t(self, p: torch.Tensor) -> torch.Tensor:
        """p fp32 [R,C] -> bf16 [C,R]: the transposed copy the input-gradient product dX = dY W reads as a K-major
        operand (opt-in, see DGRAD_TRANSPOSED_W)."""
        return self._get(("bf16_t", id(p), p.device.index), (p,), lambda: ops.cast_to_bf16_transposed(p.detach()))

    def qkv_t(self, wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
        """[w_qs; w_ks; w_vs]^T as one bf16 [D, 3*H*dk] matrix (dX of the fused QKV projection)."""
        def build():
            n = wq.shape[0]
            buf = torch.empty((wq.shape[1], 3 * n), device=wq.device, dtype=BF16)
            for i, w in enumerate((wq, wk, wv)):
                ops.cast_to_bf16_transposed(w.detach(), out=buf[:, i * n:(i + 1) * n])
            return buf
        return self._get(("qkv_t", id(wq), id(wk), id(wv), wq.device.index), (wq, wk, wv), build)

    def f32_padded(self, p: torch.Tensor, pad: int) -> torch.Tensor:
        def build():
            if pad == 0:
                return p.detach().contiguous()
            buf = torch.zeros(p.numel() + pad, device=p.device, dtype=F32)
            buf[: p.numel()].copy_(p.detach())
            return buf
        return self._get(("f32pad", id(p), p.device.index, pad), (p,), build)


CACHE = WeightCache()


def invalidate_weight_cache() -> None:
    CACHE.invalidate()


def _wgrad_split(n_out: int, k_out: int, m_red: int, n_units: int = 74) -> int:
    """split-K factor for a weight gradient [n_out, k_out] reduced over m_red rows.  The persistent GEMM runs on
    `n_units` CTA pairs (74 on a B200); work items = output tiles x splits are executed in waves of n_units, so the
    smallest split whose last wave is >= 95 % full is chosen (e.g. 64 tiles -> 8 splits = 512 items = 6.92 waves),
    keeping >= 4 k-blocks of 64 rows per split."""
    if k_out > 128 and n_out > 128:
        tiles = ((n_out + 255) // 256) * ((k_out + 255) // 256)    # CTA-pair kernel: 256 x 256 tiles
    else:
        bn = 256 if k_out > 128 else (128 if k_out > 64 else 64)
        tiles = ((n_out + 127) // 128) * ((k_out + bn - 1) // bn)
        n_units = 148
    num_kb = (m_red + 63) // 64
    max_split = max(1, min(32, num_kb // 4))
    best, best_eff = 1, 0.0
    for s in range(1, max_split + 1):
        items = tiles * s
        eff = items / (((items + n_units - 1) // n_units) * n_units)
        if eff >= 0.95:
            return s
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
    return best


# dX = dY W reads the row-major bf16 weight as an MN-major B operand.  DGRAD_TRANSPOSED_W = True (or the environment
# variable LSTC_DGRAD_TRANSPOSED_W=1) switches to a cached transposed copy (K-major B, the forward products' layout):
# alone, those products run 1 - 5 % faster (tools/gemm_ab.py), but the whole train step does not (power-capped, and the
# extra transposed casts of every weight version cost what the products gain), so it is opt-in.
DGRAD_TRANSPOSED_W = os.environ.get("LSTC_DGRAD_TRANSPOSED_W", "0") == "1"


def _dgrad(dy: torch.Tensor, w: torch.Tensor, **epilogue) -> torch.Tensor:
    """dX[m,k] = sum_n dy[m,n] w[n,k] for an nn.Linear weight w [n,k] (fp32 parameter), fused epilogue passed through."""
    if DGRAD_TRANSPOSED_W:
        return ops.gemm(dy, CACHE.bf16_t(w), **epilogue)
    return ops.gemm(dy, CACHE.bf16(w), b_mn=True, **epilogue)


def _wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW[n,k] = sum_m dy[m,n] x[m,k]  (both operands consumed MN-major, no transposes)."""
    return ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=F32,
                    split_k=_wgrad_split(dy.shape[1], x.shape[1], dy.shape[0]))


# --------------------------------------------------------------------------------------------------
class ToBF16Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_to_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return ops.cast_to_f32(g.contiguous()) if g.dtype == BF16 else g


class ToF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_to_f32(x)

    @staticmethod
    def backward(ctx, g):
        return ops.cast_to_bf16(g.contiguous()) if g.dtype == F32 else g


class ClsPrependFn(torch.autograd.Function):
    """x [W,L0,D] (fp32|bf16) -> bf16 [W,L0+1,D]."""

    @staticmethod
    def forward(ctx, x, cls_token, pos_enc, drop):
        cls = cls_token.reshape(-1) if cls_token is not None else None
        pos = pos_enc.reshape(pos_enc.shape[-2], pos_enc.shape[-1]) if pos_enc is not None else None
        out = ops.cls_prepend_fwd(x.contiguous(), cls, pos, drop if pos is not None else NO_DROPOUT)
        ctx.drop = drop if pos is not None else NO_DROPOUT
        ctx.x_dtype = x.dtype
        ctx.cls_shape = None if cls_token is None else cls_token.shape
        ctx.pos_shape = None if pos_enc is None else pos_enc.shape
        return out

    @staticmethod
    def backward(ctx, g):
        need_dx = ctx.needs_input_grad[0]
        cls_learned = ctx.cls_shape is not None
        need_dpos = ctx.pos_shape is not None and ctx.needs_input_grad[2]
        dx, dcls, dpos = ops.cls_prepend_bwd(g.contiguous(), cls_learned, need_dx, need_dpos, ctx.drop)
        if dx is not None and ctx.x_dtype == BF16:
            dx = ops.cast_to_bf16(dx)
        if dcls is not None:
            dcls = dcls.reshape(ctx.cls_shape)
        if dpos is not None:
            full = torch.zeros(ctx.pos_shape, device=g.device, dtype=F32)
            full.reshape(ctx.pos_shape[-2], ctx.pos_shape[-1])[: dpos.shape[0]].copy_(dpos)
            dpos = full
        return dx, dcls, dpos, None


class LayerNormFn(torch.autograd.Function):
    """Stand-alone LayerNorm (input_layerNorm): x fp32|bf16 [...,D] -> out_dtype."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_bf16):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, 1e-6, BF16 if out_bf16 else F32)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, _, dgamma, dbeta = ops.layernorm_bwd(g.contiguous(), x, weight, mean, rstd)
        return dx, dgamma, dbeta, None


# --------------------------------------------------------------------------------------------------
@dataclass
class MHAConfig:
    n_head: int
    d_k: int
    layer_norm: bool
    rel_mode: int          # 0 none, 1 = 3-D index sliced to [:L-1,:L-1], 2 = 2-D index (needs L-1 == ws*ws)
    attn_drop: tuple
    fc_drop: tuple
    want_attn: bool = False


class MHABlockFn(torch.autograd.Function):
    """Self-attention block on bf16 activations x [W,L,D] -> (out bf16 [W,L,D], attn fp32 [W,H,L,L] | None)."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, wfc, ln_w, ln_b, table, index, cfg: MHAConfig):
        W, L, D = x.shape
        H, dk = cfg.n_head, cfg.d_k
        x2 = x.contiguous().view(W * L, D)
        wqkv = CACHE.qkv(wq, wk, wv)
        wfc16 = CACHE.bf16(wfc)
        qkv = ops.gemm(x2, wqkv)
        bias = None
        if cfg.rel_mode:
            if cfg.rel_mode == 2 and index.shape[-1] != L - 1:
                raise RuntimeError(f"relative_pe_2D needs window_size^2 == {L - 1} tokens, index is "
                                   f"{tuple(index.shape)} (models/MultiHeadAttention.py:113-117)")
            if index.shape[-1] < L - 1:
                raise RuntimeError(f"relative position index {tuple(index.shape)} too small for {L - 1} tokens")
            bias = ops.relbias_gather(table.detach(), index, L)
        scale = 1.0 / (dk ** 0.5)
        o, probs = ops.attn_fwd(qkv, W, L, H, dk, bias, scale, cfg.attn_drop, cfg.want_attn)
        y1 = ops.gemm(o, wfc16, residual=x2, dropout=cfg.fc_drop)
        if cfg.layer_norm:
            out, mean, rstd = ops.layernorm_fwd(y1, ln_w.detach(), ln_b.detach(), 1e-6, BF16)
        else:
            out, mean, rstd = y1, None, None
        ctx.cfg, ctx.dims, ctx.scale = cfg, (W, L, D), scale
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, qkv, o, y1 if cfg.layer_norm else None, mean, rstd, wq, wk, wv, wfc, ln_w, table,
                              index)
        if probs is not None:
            ctx.mark_non_differentiable(probs)
        return out.view(W, L, D), probs

    @staticmethod
    def backward(ctx, g, _gprobs):
        x2, qkv, o, y1, mean, rstd, wq, wk, wv, wfc, ln_w, table, index = ctx.saved_tensors
        cfg: MHAConfig = ctx.cfg
        W, L, D = ctx.dims
        H, dk = cfg.n_head, cfg.d_k
        HD = H * dk
        g2 = g.contiguous().view(W * L, D)
        dlnw = dlnb = None
        if cfg.layer_norm:
            gy1, gy1d, dlnw, dlnb = ops.layernorm_bwd(g2, y1, ln_w.detach(), mean, rstd, cfg.fc_drop)
        else:
            gy1 = g2
            gy1d = ops.dropout_apply(g2, cfg.fc_drop) if cfg.fc_drop[0] > 0 else None
        gfc = gy1d if gy1d is not None else gy1
        dwfc = _wgrad(gfc, o)                                   # [D, HD]
        do = _dgrad(gfc, wfc)                                   # [M, HD]
        bias = ops.relbias_gather(table.detach(), index, L) if ctx.has_bias else None
        need_dtable = ctx.has_bias and ctx.needs_input_grad[7]
        dqkv, dbias = ops.attn_bwd(qkv, do, W, L, H, dk, bias, ctx.scale, cfg.attn_drop, need_dtable)
        dwqkv = _wgrad(dqkv, x2)                                # [3HD, D]
        gx = None
        if ctx.needs_input_grad[0]:
            if DGRAD_TRANSPOSED_W:
                gx = ops.gemm(dqkv, CACHE.qkv_t(wq, wk, wv), residual=gy1).view(W, L, D)
            else:
                gx = ops.gemm(dqkv, CACHE.qkv(wq, wk, wv), b_mn=True, residual=gy1).view(W, L, D)
        dtable = ops.relbias_scatter(dbias, index, table.shape[0]) if need_dtable else None
        return (gx, dwqkv[:HD], dwqkv[HD:2 * HD], dwqkv[2 * HD:], dwfc, dlnw, dlnb, dtable, None, None)


class MHAClsFn(torch.autograd.Function):
    """Last-layer self-attention block restricted to the CLS query (opt-in fast path, exact dead-work elimination):
    x bf16 [W,L,D] -> out bf16 [W,D] = row 0 of what MHABlockFn would return.  K and V are still projected for every
    token; Q, the attention, the out-projection, the residual and the LayerNorm run on the W CLS rows only."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, wfc, ln_w, ln_b, table, cfg: MHAConfig):
        W, L, D = x.shape
        H, dk = cfg.n_head, cfg.d_k
        HD = H * dk
        x = x.contiguous()
        x2 = x.view(W * L, D)
        xc = x.view(W, L * D)[:, :D]                      # CLS rows, row pitch L*D
        ctx.table_shape = None if table is None else tuple(table.shape)
        wqkv = CACHE.qkv(wq, wk, wv)
        kv = ops.gemm(x2, wqkv[HD:])                      # [M, 2HD]
        qc = ops.gemm(xc, wqkv[:HD])                      # [W, HD]
        scale = 1.0 / (dk ** 0.5)
        oc = ops.attn_cls_fwd(qc, kv, W, L, H, dk, scale, cfg.attn_drop)
        y1 = ops.gemm(oc, CACHE.bf16(wfc), residual=xc, dropout=cfg.fc_drop)      # [W, D]
        if cfg.layer_norm:
            out, mean, rstd = ops.layernorm_fwd(y1, ln_w.detach(), ln_b.detach(), 1e-6, BF16)
        else:
            out, mean, rstd = y1, None, None
        ctx.cfg, ctx.dims, ctx.scale = cfg, (W, L, D), scale
        ctx.save_for_backward(x, kv, qc, oc, y1 if cfg.layer_norm else None, mean, rstd, wq, wk, wv, wfc, ln_w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, kv, qc, oc, y1, mean, rstd, wq, wk, wv, wfc, ln_w = ctx.saved_tensors
        cfg: MHAConfig = ctx.cfg
        W, L, D = ctx.dims
        H, dk = cfg.n_head, cfg.d_k
        HD = H * dk
        x2 = x.view(W * L, D)
        xc = x.view(W, L * D)[:, :D]
        g2 = g.contiguous().view(W, D)
        dlnw = dlnb = None
        if cfg.layer_norm:
            gy1, gy1d, dlnw, dlnb = ops.layernorm_bwd(g2, y1, ln_w.detach(), mean, rstd, cfg.fc_drop)
        else:
            gy1 = g2
            gy1d = ops.dropout_apply(g2, cfg.fc_drop) if cfg.fc_drop[0] > 0 else None
        gfc = gy1d if gy1d is not None else gy1
        dwfc = _wgrad(gfc, oc)
        doc = ops.gemm(gfc, CACHE.bf16(wfc), b_mn=True)                          # [W, HD]
        dqc, dkv = ops.attn_cls_bwd(qc, kv, doc, W, L, H, dk, ctx.scale, cfg.attn_drop)
        dwq = _wgrad(dqc, xc)                                                    # [HD, D]
        dwkv = _wgrad(dkv, x2)                                                   # [2HD, D]
        gx = None
        if ctx.needs_input_grad[0]:
            wqkv = CACHE.qkv(wq, wk, wv)
            gx = ops.gemm(dkv, wqkv[HD:], b_mn=True)                             # [M, D]
            gcls = ops.gemm(dqc, wqkv[:HD], b_mn=True, residual=gy1)             # [W, D]: query + residual paths
            ops.add_rows_(gx.view(W, L * D)[:, :D], gcls)
            gx = gx.view(W, L, D)
        # the CLS row of the bias is zero, so the table's gradient is exactly zero here — returned as zeros (not None)
        # because the reference's autograd yields zeros and Adagrad's weight decay still touches the parameter
        dtable = None
        if ctx.table_shape is not None and ctx.needs_input_grad[7]:
            dtable = torch.zeros(ctx.table_shape, device=g.device, dtype=F32)
        return gx, dwq, dwkv[:HD], dwkv[HD:], dwfc, dlnw, dlnb, dtable, None


@dataclass
class FFNConfig:
    layer_norm: bool
    drop: tuple
    out_f32: bool = False   # last block of the encoder: the LayerNorm writes the caller's fp32 output directly


class FFNBlockFn(torch.autograd.Function):
    """x bf16 [W,L,D] -> bf16 [W,L,D]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ln_w, ln_b, cfg: FFNConfig):
        shape = x.shape
        D = shape[-1]
        x2 = x.contiguous().view(-1, D)
        Dh = w1.shape[0]
        pad = (-Dh) % 8
        w1_16 = CACHE.bf16(w1, pad_rows=pad)
        w2_16 = CACHE.bf16(w2, pad_cols=pad)
        b1p = CACHE.f32_padded(b1, pad)
        h = ops.gemm(x2, w1_16, bias=b1p, relu=True)                                    # [M, Dh+pad]
        y2 = ops.gemm(h, w2_16, bias=b2.detach().contiguous(), dropout=cfg.drop, residual=x2)
        if cfg.layer_norm:
            out, mean, rstd = ops.layernorm_fwd(y2, ln_w.detach(), ln_b.detach(), 1e-6, F32 if cfg.out_f32 else BF16)
        else:
            out, mean, rstd = y2, None, None
        ctx.cfg, ctx.shape, ctx.pad = cfg, shape, pad
        ctx.save_for_backward(x2, h, y2 if cfg.layer_norm else None, mean, rstd, w1, w2, ln_w)
        return out.view(shape)

    @staticmethod
    def backward(ctx, g):
        x2, h, y2, mean, rstd, w1, w2, ln_w = ctx.saved_tensors
        cfg: FFNConfig = ctx.cfg
        pad = ctx.pad
        Dh = w1.shape[0]
        D = x2.shape[1]
        g2 = g.contiguous().view(-1, D)
        dlnw = dlnb = None
        if cfg.layer_norm:
            # the LayerNorm backward also yields the column sums of the gradient entering w_2 (= d b_2)
            gy2, gy2d, dlnw, dlnb, db2 = ops.layernorm_bwd(g2, y2, ln_w.detach(), mean, rstd, cfg.drop, want_dxsum=True)
            gin = gy2d if gy2d is not None else gy2
        else:
            gy2 = g2
            gy2d = ops.dropout_apply(g2, cfg.drop) if cfg.drop[0] > 0 else None
            gin = gy2d if gy2d is not None else gy2
            db2 = ops.colsum(gin)
        dw2 = _wgrad(gin, h)                                                  # [D, Dh+pad]
        # d b_1 = column sums of dh: accumulated by the product's epilogue from the staged bf16 tile where it can
        fused = ops.gemm_fuses_colsum(gin.shape[0], Dh + pad)
        db1 = torch.zeros(Dh + pad, device=gin.device, dtype=F32) if fused else None
        if pad == 0:
            dh = _dgrad(gin, w2, relu_mask=h, colsum_out=db1)                 # [M, Dh]
        else:
            dh = ops.gemm(gin, CACHE.bf16(w2, pad_cols=pad), b_mn=True, relu_mask=h, colsum_out=db1)   # [M, Dh+pad]
        if not fused:
            db1 = ops.colsum(dh)
        dw1 = _wgrad(dh, x2)                                                  # [Dh+pad, D]
        gx = None
        if ctx.needs_input_grad[0]:
            if pad == 0:
                gx = _dgrad(dh, w1, residual=gy2).view(ctx.shape)
            else:
                gx = ops.gemm(dh, CACHE.bf16(w1, pad_rows=pad), b_mn=True, residual=gy2).view(ctx.shape)
        if pad:
            dw2, dw1, db1 = dw2[:, :Dh].contiguous(), dw1[:Dh], db1[:Dh]
        return gx, dw1, db1, dw2, db2, dlnw, dlnb, None


# --------------------------------------------------------------------------------------------------
@dataclass
class HeadConfig:
    sigmoid: bool
    drop1: tuple
    drop2: tuple


class HeadFn(torch.autograd.Function):
    """3-layer MLP head: x [n,D] (fp32|bf16, row-strided ok) -> fp32 [n,C] probabilities / scores."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, cfg: HeadConfig):
        if x.dtype == F32:
            x16 = ops.cast_to_bf16(x)
        else:
            x16 = x if (x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 8 == 0) else x.contiguous()
        h1 = ops.gemm(x16, CACHE.bf16(w1), bias=b1.detach().contiguous(), relu=True, dropout=cfg.drop1)
        h2, out = ops.head_tail_fwd(h1, w2.detach(), b2.detach(), w3.detach(), b3.detach(), cfg.sigmoid, cfg.drop2)
        ctx.cfg, ctx.x_dtype = cfg, x.dtype
        ctx.save_for_backward(x16, h1, h2, out, w1, w2, w3)
        return out

    @staticmethod
    def backward(ctx, dout):
        x16, h1, h2, out, w1, w2, w3 = ctx.saved_tensors
        cfg: HeadConfig = ctx.cfg
        dh2, dw3, db3 = ops.head_tail_bwd(dout.contiguous().float(), out, h2, w3.detach(), cfg.sigmoid, cfg.drop2)
        db2 = ops.colsum(dh2)
        dw2 = ops.gemm(dh2, h1, a_mn=True, b_mn=True, out_dtype=F32)          # [32, K1]
        # h1 > 0 <=> kept by dropout and positive pre-activation; the dropout epilogue re-applies the 1/(1-p) scale
        dh1 = ops.gemm(dh2, CACHE.bf16(w2), b_mn=True, relu_mask=h1, dropout=cfg.drop1)
        db1 = ops.colsum(dh1)
        dw1 = _wgrad(dh1, x16)                                                # [K1, D]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dh1, CACHE.bf16(w1), b_mn=True, out_dtype=ctx.x_dtype)
        return dx, dw1, db1, dw2, db2, dw3, db3, None


# --------------------------------------------------------------------------------------------------
class MILLossFn(torch.autograd.Function):
    """-> (loss, err, spar) 0-dim fp32 tensors; only `loss` is differentiable (err / spar are logged values)."""

    @staticmethod
    def forward(ctx, y_pred, B, P, T, topk, lambda1, spar_start):
        y = y_pred.reshape(-1)
        if y.dtype != F32:
            y = y.float()
        out3, top_idx, dscores = ops.mil_loss(y.contiguous(), B, P, T, topk, lambda1, spar_start,
                                              need_grad=ctx.needs_input_grad[0])
        ctx.shape = y_pred.shape
        ctx.save_for_backward(dscores)
        loss, err, spar = out3[0], out3[1], out3[2]
        ctx.mark_non_differentiable(err, spar, top_idx)
        return loss, err, spar, top_idx

    @staticmethod
    def backward(ctx, g, _ge, _gs, _gi):
        (dscores,) = ctx.saved_tensors
        grad = ops.scale_by_device_scalar(dscores, g.reshape(1).contiguous()).view(ctx.shape)
        return grad, None, None, None, None, None, None


class SoftCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels):
        out1, dprobs = ops.soft_ce_loss(outputs.float(), labels.float(), need_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(dprobs)
        return out1[0]

    @staticmethod
    def backward(ctx, g):
        (dprobs,) = ctx.saved_tensors
        return ops.scale_by_device_scalar(dprobs, g.reshape(1).contiguous()), None


class BCEFn(torch.autograd.Function):
    """scores [n_parts*T] (mean over T taken inside), labels [n_parts,2]."""

    @staticmethod
    def forward(ctx, scores, labels, T, w_normal, w_abnormal):
        out1, ds = ops.bce_loss(scores.reshape(-1).float(), labels.reshape(-1, 2).float(), T, w_normal, w_abnormal,
                                need_grad=ctx.needs_input_grad[0])
        ctx.shape = scores.shape
        ctx.save_for_backward(ds)
        return out1[0]

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        return ops.scale_by_device_scalar(ds, g.reshape(1).contiguous()).view(ctx.shape), None, None, None, None
