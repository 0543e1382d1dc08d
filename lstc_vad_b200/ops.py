"""Tensor-level wrappers over the C-ABI (no autograd here; see functional.py).

torch is used only for device memory and streams: every wrapper checks that its operands are CUDA tensors on
the current device, passes raw device pointers + the current stream to the library and raises on failure.
There is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32

Dropout = Tuple[float, int, int]  # (p, seed, offset)
NO_DROPOUT: Dropout = (0.0, 0, 0)


class _LaunchCounter:
    """Counts kernel launches issued through this module (bench.py reports it as `gpu_launches`)."""

    def __init__(self):
        self.count = 0

    def reset(self):
        self.count = 0

    def add(self, n: int = 1):
        self.count += n


class _GemmProfiler:
    """Optional CUDA-event bracket around every GEMM launch (on the launching stream) — used by bench.py to
    measure the tensor-core kernel live inside the timed region."""

    def __init__(self):
        self.enabled = False
        self.records = []

    def enable(self):
        self.enabled, self.records = True, []

    def disable(self):
        self.enabled = False

    def collect(self):
        torch.cuda.synchronize()
        out = [dict(kind=k, flops=f, ms=e0.elapsed_time(e1), shape=shape) for k, f, shape, e0, e1 in self.records]
        self.records = []
        return out


LAUNCHES = _LaunchCounter()
PROFILE = _GemmProfiler()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"lstc_vad_b200: `{name}` must be a CUDA tensor (there is no CPU fallback), got "
                           f"{t.device if isinstance(t, torch.Tensor) else type(t)}")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"lstc_vad_b200: `{name}` lives on {t.device} but the current device is "
                           f"cuda:{torch.cuda.current_device()}")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"lstc_vad_b200: `{name}` must be {dtype}, got {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _rowmajor2d(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise RuntimeError(f"lstc_vad_b200: `{name}` must be a 2-D row-major matrix, got shape {tuple(t.shape)} "
                           f"strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# ------------------------------------------------------------------------------------------ GEMM
def gemm_fuses_colsum(M: int, N: int, ldc: Optional[int] = None, out_dtype=BF16) -> bool:
    """True when `gemm(..., colsum_out=...)` can add the column sums of its bf16 output in the epilogue."""
    return bool(_lib.load().lstc_gemm_bf16_fuses_colsum(int(M), int(N), int(N if ldc is None else ldc), int(out_dtype == F32)))


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out_dtype=BF16,
         bias: Optional[torch.Tensor] = None, relu: bool = False, relu_mask: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, dropout: Dropout = NO_DROPOUT, split_k: int = 1,
         out: Optional[torch.Tensor] = None, accumulate: bool = False,
         colsum_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] = epilogue(A @ B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn); bf16.
    colsum_out (fp32 [N]): += column sums of the stored bf16 C, fused into the epilogue (see gemm_fuses_colsum)."""
    lib = _lib.load()
    _cuda(a, "a", BF16)
    _cuda(b, "b", BF16)
    lda, ldb = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    (K, M) = a.shape if a_mn else a.shape[::-1]
    (Kb, N) = b.shape if b_mn else b.shape[::-1]
    if K != Kb:
        raise RuntimeError(f"lstc_vad_b200.gemm: reduction dims differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    else:
        _cuda(out, "out")
        if tuple(out.shape) != (M, N):
            raise RuntimeError(f"lstc_vad_b200.gemm: out has shape {tuple(out.shape)}, expected {(M, N)}")
    ldc = _rowmajor2d(out, "out")
    if out.dtype not in (BF16, F32):
        raise RuntimeError("lstc_vad_b200.gemm: out must be bf16 or fp32")
    if bias is not None:
        _cuda(bias, "bias", F32)
        if bias.numel() != N or not bias.is_contiguous():
            raise RuntimeError("lstc_vad_b200.gemm: bias must be a contiguous fp32 [N] vector")
    ld_mask = ld_res = 0
    if relu_mask is not None:
        _cuda(relu_mask, "relu_mask", BF16)
        ld_mask = _rowmajor2d(relu_mask, "relu_mask")
    if residual is not None:
        _cuda(residual, "residual", BF16)
        ld_res = _rowmajor2d(residual, "residual")
    if colsum_out is not None:
        _cuda(colsum_out, "colsum_out", F32)
        if colsum_out.numel() != N or not colsum_out.is_contiguous():
            raise RuntimeError("lstc_vad_b200.gemm: colsum_out must be a contiguous fp32 [N] vector")
    p, seed, off = dropout
    prof = PROFILE.enabled
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    st = lib.lstc_gemm_bf16(_p(a), lda, int(a_mn), _p(b), ldb, int(b_mn), M, N, K, _p(out), ldc,
                            int(out.dtype == F32), _p(bias), int(relu), _p(relu_mask), ld_mask, _p(residual), ld_res,
                            float(p), int(seed), int(off), int(split_k), int(accumulate), _p(colsum_out), _stream())
    if prof:
        e1.record()
        kind = ("mn" if a_mn else "k") + "-" + ("mn" if b_mn else "k")
        PROFILE.records.append((kind, 2.0 * M * N * K, (M, N, K), e0, e1))
    _lib.check(st, "lstc_gemm_bf16")
    LAUNCHES.add(1)
    return out


# ------------------------------------------------------------------------------------------ attention
def relbias_gather(table: torch.Tensor, index: torch.Tensor, L: int) -> torch.Tensor:
    """table [T,H] fp32, index [n,n] int64 -> dense [H,L,L] fp32 with a zero CLS row/column."""
    lib = _lib.load()
    _cuda(table, "table", F32)
    _cuda(index, "index", torch.int64)
    table = table.contiguous()
    index = index.contiguous()
    T, H = table.shape
    dense = torch.empty((H, L, L), device=table.device, dtype=F32)
    st = lib.lstc_relbias_gather(_p(table), _p(index), index.shape[-1], L, H, T, _p(dense), _stream())
    _lib.check(st, "lstc_relbias_gather")
    LAUNCHES.add(1)
    return dense


def relbias_scatter(ddense: torch.Tensor, index: torch.Tensor, T: int) -> torch.Tensor:
    lib = _lib.load()
    _cuda(ddense, "ddense", F32)
    _cuda(index, "index", torch.int64)
    H, L, _ = ddense.shape
    index = index.contiguous()
    dtable = torch.empty((T, H), device=ddense.device, dtype=F32)
    st = lib.lstc_relbias_scatter(_p(ddense.contiguous()), _p(index), index.shape[-1], L, H, T, _p(dtable), _stream())
    _lib.check(st, "lstc_relbias_scatter")
    LAUNCHES.add(1)
    return dtable


def attn_fwd(qkv: torch.Tensor, W: int, L: int, H: int, dk: int, bias: Optional[torch.Tensor], scale: float,
             dropout: Dropout = NO_DROPOUT, return_probs: bool = False):
    """qkv bf16 [W*L, 3*H*dk] -> (out bf16 [W*L, H*dk], probs fp32 [W,H,L,L] | None)."""
    lib = _lib.load()
    _cuda(qkv, "qkv", BF16)
    ld = _rowmajor2d(qkv, "qkv")
    if qkv.shape[0] != W * L or qkv.shape[1] != 3 * H * dk:
        raise RuntimeError(f"lstc_vad_b200.attn_fwd: qkv shape {tuple(qkv.shape)} != {(W * L, 3 * H * dk)}")
    if bias is not None:
        _cuda(bias, "bias", F32)
        if tuple(bias.shape) != (H, L, L) or not bias.is_contiguous():
            raise RuntimeError("lstc_vad_b200.attn_fwd: bias must be contiguous fp32 [H,L,L]")
    out = torch.empty((W * L, H * dk), device=qkv.device, dtype=BF16)
    probs = torch.empty((W, H, L, L), device=qkv.device, dtype=F32) if return_probs else None
    p, seed, off = dropout
    st = lib.lstc_attn_fwd(_p(qkv), ld, W, L, H, dk, _p(bias), float(scale), float(p), int(seed), int(off), _p(out),
                           H * dk, _p(probs), _stream())
    _lib.check(st, "lstc_attn_fwd")
    LAUNCHES.add(1)
    return out, probs


def attn_bwd(qkv: torch.Tensor, dout: torch.Tensor, W: int, L: int, H: int, dk: int, bias: Optional[torch.Tensor],
             scale: float, dropout: Dropout = NO_DROPOUT, need_dbias: bool = False):
    """-> (dqkv bf16 [W*L, 3*H*dk], dbias fp32 [H,L,L] | None)."""
    lib = _lib.load()
    _cuda(qkv, "qkv", BF16)
    _cuda(dout, "dout", BF16)
    ld, ldo = _rowmajor2d(qkv, "qkv"), _rowmajor2d(dout, "dout")
    if tuple(dout.shape) != (W * L, H * dk):
        raise RuntimeError(f"lstc_vad_b200.attn_bwd: dout shape {tuple(dout.shape)} != {(W * L, H * dk)}")
    dqkv = torch.empty((W * L, 3 * H * dk), device=qkv.device, dtype=BF16)
    dbias = torch.empty((H, L, L), device=qkv.device, dtype=F32) if need_dbias else None
    p, seed, off = dropout
    st = lib.lstc_attn_bwd(_p(qkv), ld, _p(dout), ldo, W, L, H, dk, _p(bias), float(scale), float(p), int(seed),
                           int(off), _p(dqkv), 3 * H * dk, _p(dbias), _stream())
    _lib.check(st, "lstc_attn_bwd")
    LAUNCHES.add(1)
    return dqkv, dbias


def attn_cls_fwd(q: torch.Tensor, kv: torch.Tensor, W: int, L: int, H: int, dk: int, scale: float,
                 dropout: Dropout = NO_DROPOUT) -> torch.Tensor:
    """CLS-query attention: q bf16 [W, H*dk], kv bf16 [W*L, 2*H*dk] (k | v column blocks) -> o bf16 [W, H*dk]."""
    lib = _lib.load()
    _cuda(q, "q", BF16)
    _cuda(kv, "kv", BF16)
    ldq, ldkv = _rowmajor2d(q, "q"), _rowmajor2d(kv, "kv")
    HD = H * dk
    if tuple(q.shape) != (W, HD) or tuple(kv.shape) != (W * L, 2 * HD):
        raise RuntimeError(f"lstc_vad_b200.attn_cls_fwd: shapes {tuple(q.shape)}, {tuple(kv.shape)}")
    out = torch.empty((W, HD), device=q.device, dtype=BF16)
    p, seed, off = dropout
    st = lib.lstc_attn_cls_fwd(_p(q), ldq, _p(kv), kv.data_ptr() + 2 * HD, ldkv, W, L, H, dk, float(scale), float(p),
                               int(seed), int(off), _p(out), HD, _stream())
    _lib.check(st, "lstc_attn_cls_fwd")
    LAUNCHES.add(1)
    return out


def attn_cls_bwd(q: torch.Tensor, kv: torch.Tensor, dout: torch.Tensor, W: int, L: int, H: int, dk: int, scale: float,
                 dropout: Dropout = NO_DROPOUT):
    """-> (dq bf16 [W, H*dk], dkv bf16 [W*L, 2*H*dk])."""
    lib = _lib.load()
    _cuda(q, "q", BF16)
    _cuda(kv, "kv", BF16)
    _cuda(dout, "dout", BF16)
    ldq, ldkv, ldo = _rowmajor2d(q, "q"), _rowmajor2d(kv, "kv"), _rowmajor2d(dout, "dout")
    HD = H * dk
    dq = torch.empty((W, HD), device=q.device, dtype=BF16)
    dkv = torch.empty((W * L, 2 * HD), device=q.device, dtype=BF16)
    p, seed, off = dropout
    st = lib.lstc_attn_cls_bwd(_p(q), ldq, _p(kv), kv.data_ptr() + 2 * HD, ldkv, _p(dout), ldo, W, L, H, dk,
                               float(scale), float(p), int(seed), int(off), _p(dq), HD, _p(dkv),
                               dkv.data_ptr() + 2 * HD, 2 * HD, _stream())
    _lib.check(st, "lstc_attn_cls_bwd")
    LAUNCHES.add(1)
    return dq, dkv


def add_rows_(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst += src for 2-D bf16 row-strided views of equal shape."""
    lib = _lib.load()
    _cuda(dst, "dst", BF16)
    _cuda(src, "src", BF16)
    ldd, lds = _rowmajor2d(dst, "dst"), _rowmajor2d(src, "src")
    if dst.shape != src.shape:
        raise RuntimeError("lstc_vad_b200.add_rows_: shape mismatch")
    st = lib.lstc_add_rows_bf16(_p(dst), ldd, _p(src), lds, dst.shape[0], dst.shape[1], _stream())
    _lib.check(st, "lstc_add_rows_bf16")
    LAUNCHES.add(1)


# ------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6, out_dtype=BF16):
    """x [..., D] bf16|fp32 -> (y, mean[rows], rstd[rows])."""
    lib = _lib.load()
    _cuda(x, "x")
    _cuda(gamma, "gamma", F32)
    _cuda(beta, "beta", F32)
    if x.dtype not in (BF16, F32) or not x.is_contiguous():
        raise RuntimeError("lstc_vad_b200.layernorm_fwd: x must be contiguous bf16 or fp32")
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    mean = torch.empty(rows, device=x.device, dtype=F32)
    rstd = torch.empty(rows, device=x.device, dtype=F32)
    st = lib.lstc_layernorm_fwd(_p(x), int(x.dtype == F32), _p(gamma), _p(beta), _p(y), int(out_dtype == F32),
                                _p(mean), _p(rstd), rows, D, float(eps), _stream())
    _lib.check(st, "lstc_layernorm_fwd")
    LAUNCHES.add(1)
    return y, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                  dropout: Dropout = NO_DROPOUT, want_dxsum: bool = False):
    """-> (dx [same dtype as x], dx_drop bf16 | None, dgamma, dbeta[, dxsum if want_dxsum])."""
    lib = _lib.load()
    _cuda(dy, "dy")
    _cuda(x, "x")
    if not dy.is_contiguous() or not x.is_contiguous():
        raise RuntimeError("lstc_vad_b200.layernorm_bwd: dy and x must be contiguous")
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty_like(x)
    p, seed, off = dropout
    dx_drop = torch.empty(x.shape, device=x.device, dtype=BF16) if p > 0 else None
    dgamma = torch.empty(D, device=x.device, dtype=F32)
    dbeta = torch.empty(D, device=x.device, dtype=F32)
    ws = torch.empty(max(1, lib.lstc_layernorm_bwd_workspace(rows, D)), device=x.device, dtype=torch.uint8)
    dxsum = torch.empty(D, device=x.device, dtype=F32) if want_dxsum else None
    st = lib.lstc_layernorm_bwd(_p(dy), int(dy.dtype == F32), _p(x), int(x.dtype == F32), _p(gamma), _p(mean),
                                _p(rstd), _p(dx), int(dx.dtype == F32), _p(dx_drop), float(p), int(seed), int(off),
                                _p(dgamma), _p(dbeta), _p(dxsum), _p(ws), rows, D, _stream())
    _lib.check(st, "lstc_layernorm_bwd")
    LAUNCHES.add(2)
    if want_dxsum:
        return dx, dx_drop, dgamma, dbeta, dxsum
    return dx, dx_drop, dgamma, dbeta


# ------------------------------------------------------------------------------------------ CLS prepend
def cls_prepend_fwd(x: torch.Tensor, cls: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
                    dropout: Dropout = NO_DROPOUT) -> torch.Tensor:
    """x [W,L0,D] fp32|bf16 -> bf16 [W,L0+1,D]; cls fp32 [D] or None (token mean); pos fp32 [>=L0+1, D]."""
    lib = _lib.load()
    _cuda(x, "x")
    if x.dim() != 3 or x.dtype not in (BF16, F32) or not x.is_contiguous():
        raise RuntimeError("lstc_vad_b200.cls_prepend_fwd: x must be contiguous [W,L0,D] bf16 or fp32")
    W, L0, D = x.shape
    if cls is not None:
        _cuda(cls, "cls", F32)
        cls = cls.contiguous()
    if pos is not None:
        _cuda(pos, "pos", F32)
        pos = pos.contiguous()
        if pos.shape[-2] < L0 + 1 or pos.shape[-1] != D:
            raise RuntimeError(f"lstc_vad_b200.cls_prepend_fwd: position table {tuple(pos.shape)} too small for "
                               f"{L0 + 1} tokens")
    out = torch.empty((W, L0 + 1, D), device=x.device, dtype=BF16)
    p, seed, off = dropout
    st = lib.lstc_cls_prepend_fwd(_p(x), int(x.dtype == F32), _p(cls), _p(pos), float(p), int(seed), int(off),
                                  _p(out), W, L0, D, _stream())
    _lib.check(st, "lstc_cls_prepend_fwd")
    LAUNCHES.add(1)
    return out


def cls_prepend_bwd(g: torch.Tensor, cls_learned: bool, need_dx: bool, need_dpos: bool,
                    dropout: Dropout = NO_DROPOUT):
    """g bf16 [W,L0+1,D] -> (dx fp32 [W,L0,D] | None, dcls fp32 [D] | None, dpos fp32 [L0+1,D] | None)."""
    lib = _lib.load()
    _cuda(g, "g", BF16)
    g = g.contiguous()
    W, L, D = g.shape
    L0 = L - 1
    dx = torch.empty((W, L0, D), device=g.device, dtype=F32) if need_dx else None
    dcls = torch.empty(D, device=g.device, dtype=F32) if cls_learned else None
    dpos = torch.empty((L, D), device=g.device, dtype=F32) if need_dpos else None
    p, seed, off = dropout
    st = lib.lstc_cls_prepend_bwd(_p(g), int(cls_learned), float(p), int(seed), int(off), _p(dx), _p(dcls), _p(dpos),
                                  W, L0, D, _stream())
    _lib.check(st, "lstc_cls_prepend_bwd")
    LAUNCHES.add(1)
    return dx, dcls, dpos


# ------------------------------------------------------------------------------------------ heads
def head_tail_fwd(h1: torch.Tensor, W2, b2, W3, b3, sigmoid: bool, dropout: Dropout = NO_DROPOUT):
    """h1 bf16 [n,K1] -> (h2 bf16 [n,32], out fp32 [n,C])."""
    lib = _lib.load()
    _cuda(h1, "h1", BF16)
    for nm, t in (("W2", W2), ("b2", b2), ("W3", W3), ("b3", b3)):
        _cuda(t, nm, F32)
    if not h1.is_contiguous() or W2.shape[0] != 32 or W3.shape[1] != 32:
        raise RuntimeError("lstc_vad_b200.head_tail_fwd: expects contiguous h1 and a 32-wide middle layer")
    n, K1 = h1.shape
    C = W3.shape[0]
    h2 = torch.empty((n, 32), device=h1.device, dtype=BF16)
    out = torch.empty((n, C), device=h1.device, dtype=F32)
    p, seed, off = dropout
    st = lib.lstc_head_tail_fwd(_p(h1), n, K1, _p(W2.contiguous()), _p(b2.contiguous()), _p(W3.contiguous()),
                                _p(b3.contiguous()), C, int(sigmoid), float(p), int(seed), int(off), _p(h2), _p(out),
                                _stream())
    _lib.check(st, "lstc_head_tail_fwd")
    LAUNCHES.add(1)
    return h2, out


def head_tail_bwd(dout: torch.Tensor, out: torch.Tensor, h2: torch.Tensor, W3: torch.Tensor, sigmoid: bool,
                  dropout: Dropout = NO_DROPOUT):
    """-> (dh2 bf16 [n,32], dW3 fp32 [C,32], db3 fp32 [C])."""
    lib = _lib.load()
    _cuda(dout, "dout", F32)
    _cuda(out, "out", F32)
    _cuda(h2, "h2", BF16)
    _cuda(W3, "W3", F32)
    n, C = out.shape
    dout = dout.contiguous()
    dh2 = torch.empty((n, 32), device=out.device, dtype=BF16)
    dW3 = torch.empty((C, 32), device=out.device, dtype=F32)
    db3 = torch.empty(C, device=out.device, dtype=F32)
    p, seed, off = dropout
    st = lib.lstc_head_tail_bwd(_p(dout), _p(out), _p(h2), _p(W3.contiguous()), n, C, int(sigmoid), float(p),
                                int(seed), int(off), _p(dh2), _p(dW3), _p(db3), _stream())
    _lib.check(st, "lstc_head_tail_bwd")
    LAUNCHES.add(1)
    return dh2, dW3, db3


# ------------------------------------------------------------------------------------------ losses
def mil_loss(scores: torch.Tensor, B: int, P: int, T: int = 1, topk: int = 1, lambda1: float = 0.01,
             spar_start: Optional[int] = None, need_grad: bool = True):
    """scores: fp32, 2*B*P*T elements, possibly a strided 1-D view (e.g. probs[:, 1]).
    -> (out3 [loss, err, spar], top_idx int32 [2B,topk], dscores (same layout as a dense copy) | None)"""
    lib = _lib.load()
    _cuda(scores, "scores", F32)
    n = 2 * B * P * T
    if scores.numel() != n:
        raise RuntimeError(f"lstc_vad_b200.mil_loss: expected {n} scores (2*B*P*T), got {scores.numel()}")
    flat = scores.reshape(-1)
    stride = flat.stride(0) if n > 1 else 1
    if stride < 1:
        flat, stride = flat.contiguous(), 1
    if spar_start is None:
        spar_start = B
    out3 = torch.empty(3, device=scores.device, dtype=F32)
    top_idx = torch.empty((2 * B, topk), device=scores.device, dtype=torch.int32)
    dscores = torch.empty(n, device=scores.device, dtype=F32) if need_grad else None
    # the gradient is always written densely (stride 1 into `dscores`)
    if stride != 1:
        flat = flat.contiguous()
    st = lib.lstc_mil_loss(_p(flat), 1, B, P, T, topk, float(lambda1), int(spar_start), _p(out3), _p(top_idx),
                           _p(dscores), _stream())
    _lib.check(st, "lstc_mil_loss")
    LAUNCHES.add(1)
    return out3, top_idx, dscores


def soft_ce_loss(probs: torch.Tensor, labels: torch.Tensor, need_grad: bool = True):
    lib = _lib.load()
    _cuda(probs, "probs", F32)
    _cuda(labels, "labels", F32)
    probs, labels = probs.contiguous(), labels.contiguous()
    n, C = probs.shape
    if tuple(labels.shape) != (n, C):
        raise RuntimeError("lstc_vad_b200.soft_ce_loss: labels must match probs' shape")
    out1 = torch.empty(1, device=probs.device, dtype=F32)
    dprobs = torch.empty_like(probs) if need_grad else None
    st = lib.lstc_soft_ce_loss(_p(probs), _p(labels), n, C, _p(out1), _p(dprobs), _stream())
    _lib.check(st, "lstc_soft_ce_loss")
    LAUNCHES.add(1)
    return out1, dprobs


def bce_loss(scores: torch.Tensor, labels: torch.Tensor, T: int, w_normal: float, w_abnormal: float,
             need_grad: bool = True):
    """scores fp32 [n_parts*T]; labels fp32 [n_parts, 2]."""
    lib = _lib.load()
    _cuda(scores, "scores", F32)
    _cuda(labels, "labels", F32)
    scores, labels = scores.contiguous(), labels.contiguous()
    n_parts = labels.numel() // 2
    if scores.numel() != n_parts * T:
        raise RuntimeError("lstc_vad_b200.bce_loss: scores must hold n_parts*T elements")
    out1 = torch.empty(1, device=scores.device, dtype=F32)
    dscores = torch.empty(scores.numel(), device=scores.device, dtype=F32) if need_grad else None
    st = lib.lstc_bce_loss(_p(scores), _p(labels), n_parts, T, float(w_normal), float(w_abnormal), _p(out1),
                           _p(dscores), _stream())
    _lib.check(st, "lstc_bce_loss")
    LAUNCHES.add(1)
    return out1, dscores


def weighted_auc(scores: torch.Tensor, pos_w: torch.Tensor, neg_w: torch.Tensor) -> torch.Tensor:
    """ROC-AUC of n scored groups, group i holding pos_w[i] positives and neg_w[i] negatives that all carry
    scores[i].  -> fp64 [3] = (auc, total positives, total negatives) on the device."""
    lib = _lib.load()
    for nm, t in (("scores", scores), ("pos_w", pos_w), ("neg_w", neg_w)):
        _cuda(t, nm, F32)
    scores, pos_w, neg_w = scores.contiguous().reshape(-1), pos_w.contiguous().reshape(-1), neg_w.contiguous().reshape(-1)
    if not (scores.numel() == pos_w.numel() == neg_w.numel()):
        raise RuntimeError("lstc_vad_b200.weighted_auc: scores / pos_w / neg_w must have the same length")
    out = torch.empty(3, device=scores.device, dtype=torch.float64)
    st = lib.lstc_weighted_auc(_p(scores), _p(pos_w), _p(neg_w), scores.numel(), _p(out), _stream())
    _lib.check(st, "lstc_weighted_auc")
    LAUNCHES.add(1)
    return out


def threshold_labels(scores: torch.Tensor, thr: float) -> torch.Tensor:
    lib = _lib.load()
    _cuda(scores, "scores", F32)
    scores = scores.contiguous()
    out = torch.empty_like(scores)
    st = lib.lstc_threshold_labels(_p(scores), float(thr), _p(out), scores.numel(), _stream())
    _lib.check(st, "lstc_threshold_labels")
    LAUNCHES.add(1)
    return out


# ------------------------------------------------------------------------------------------ utilities
def cast_to_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _cuda(x, "x", F32)
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    st = lib.lstc_cast_f32_to_bf16(_p(x), _p(out), x.numel(), _stream())
    _lib.check(st, "lstc_cast_f32_to_bf16")
    LAUNCHES.add(1)
    return out


def cast_to_bf16_transposed(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [R, C] -> bf16 [C, R] (out may be a row-major view with a wider pitch)."""
    lib = _lib.load()
    _cuda(x, "x", F32)
    ld_src = _rowmajor2d(x, "x")
    R, C = x.shape
    if out is None:
        out = torch.empty((C, R), device=x.device, dtype=BF16)
    else:
        _cuda(out, "out", BF16)
        if tuple(out.shape) != (C, R):
            raise RuntimeError(f"lstc_vad_b200.cast_to_bf16_transposed: out has shape {tuple(out.shape)}, expected {(C, R)}")
    st = lib.lstc_cast_f32_to_bf16_transposed(_p(x), ld_src, _p(out), _rowmajor2d(out, "out"), R, C, _stream())
    _lib.check(st, "lstc_cast_f32_to_bf16_transposed")
    LAUNCHES.add(1)
    return out


def cast_to_f32(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _cuda(x, "x", BF16)
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=F32)
    st = lib.lstc_cast_bf16_to_f32(_p(x), _p(out), x.numel(), _stream())
    _lib.check(st, "lstc_cast_bf16_to_f32")
    LAUNCHES.add(1)
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """x bf16 [rows, cols] -> fp32 [cols]."""
    lib = _lib.load()
    _cuda(x, "x", BF16)
    ld = _rowmajor2d(x, "x")
    rows, cols = x.shape
    out = torch.empty(cols, device=x.device, dtype=F32)
    ws = torch.empty(max(1, lib.lstc_colsum_workspace(rows, cols)), device=x.device, dtype=torch.uint8)
    st = lib.lstc_colsum_bf16(_p(x), rows, cols, ld, _p(out), _p(ws), _stream())
    _lib.check(st, "lstc_colsum_bf16")
    LAUNCHES.add(2)
    return out


def dropout_apply(x: torch.Tensor, dropout: Dropout) -> torch.Tensor:
    lib = _lib.load()
    _cuda(x, "x", BF16)
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty_like(x)
    p, seed, off = dropout
    st = lib.lstc_dropout_apply_bf16(_p(x), _p(y), rows, cols, float(p), int(seed), int(off), _stream())
    _lib.check(st, "lstc_dropout_apply_bf16")
    LAUNCHES.add(1)
    return y


def dropout_mask(rows: int, cols: int, dropout: Dropout, device=None) -> torch.Tensor:
    """The exact keep-mask (uint8 [rows, cols]) the fused kernels generate for (p, seed, offset)."""
    lib = _lib.load()
    device = device or torch.device("cuda", torch.cuda.current_device())
    mask = torch.empty((rows, cols), device=device, dtype=torch.uint8)
    p, seed, off = dropout
    st = lib.lstc_dropout_mask(_p(mask), rows, cols, float(p), int(seed), int(off), _stream())
    _lib.check(st, "lstc_dropout_mask")
    LAUNCHES.add(1)
    return mask


def scale_by_device_scalar(src: torch.Tensor, scalar: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _cuda(src, "src", F32)
    _cuda(scalar, "scalar", F32)
    src = src.contiguous()
    dst = torch.empty_like(src)
    st = lib.lstc_scale_by_device_scalar(_p(src), _p(scalar), _p(dst), src.numel(), _stream())
    _lib.check(st, "lstc_scale_by_device_scalar")
    LAUNCHES.add(1)
    return dst


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """src [n, ...] (contiguous, row size a multiple of 16 bytes), idx int64 [m] on the device -> [m, ...]."""
    lib = _lib.load()
    _cuda(src, "src")
    _cuda(idx, "idx", torch.int64)
    src, idx = src.contiguous(), idx.contiguous()
    row_bytes = src[0].numel() * src.element_size()
    out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    st = lib.lstc_gather_rows(_p(src), row_bytes, _p(idx), idx.numel(), _p(out), _stream())
    _lib.check(st, "lstc_gather_rows")
    LAUNCHES.add(1)
    return out


def segment_mean(feats: torch.Tensor, bounds: torch.Tensor, l2norm: bool = False) -> torch.Tensor:
    """feats fp32 [n_clips, n_patch, D], bounds int32 [n_bins+1] -> fp32 [n_bins, n_patch, D] (bin means, empty bin =
    its first clip; optional per-token L2 normalisation)."""
    lib = _lib.load()
    _cuda(feats, "feats", F32)
    _cuda(bounds, "bounds", torch.int32)
    feats, bounds = feats.contiguous(), bounds.contiguous()
    n_clips, n_patch, D = feats.shape
    n_bins = bounds.numel() - 1
    out = torch.empty((n_bins, n_patch, D), device=feats.device, dtype=F32)
    st = lib.lstc_segment_mean(_p(feats), _p(bounds), n_bins, n_clips, n_patch, D, int(l2norm), _p(out), _stream())
    _lib.check(st, "lstc_segment_mean")
    LAUNCHES.add(1)
    return out


def grad_clip_coef(grads, max_norm: float) -> torch.Tensor:
    """Device scalar min(1, max_norm / (||grads||_2 + 1e-6)) over a list of fp32 tensors — clip_grad_norm_ without a
    host sync (the coefficient is consumed by adagrad_step)."""
    lib = _lib.load()
    dev = grads[0].device
    acc = torch.zeros(2, device=dev, dtype=F32)
    for g in grads:
        _cuda(g, "grad", F32)
        g = g.contiguous()
        st = lib.lstc_sumsq_accumulate(_p(g), g.numel(), _p(acc), _stream())
        _lib.check(st, "lstc_sumsq_accumulate")
        LAUNCHES.add(1)
    st = lib.lstc_clip_coef(_p(acc), float(max_norm), acc.data_ptr() + 4, _stream())
    _lib.check(st, "lstc_clip_coef")
    LAUNCHES.add(1)
    return acc[1:2]


def adagrad_step(param: torch.Tensor, grad: torch.Tensor, state_sum: torch.Tensor, lr: float, weight_decay: float,
                 eps: float = 1e-10, grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None) -> None:
    lib = _lib.load()
    for nm, t in (("param", param), ("grad", grad), ("state_sum", state_sum)):
        _cuda(t, nm, F32)
        if not t.is_contiguous():
            raise RuntimeError(f"lstc_vad_b200.adagrad_step: `{nm}` must be contiguous")
    st = lib.lstc_adagrad_step(_p(param), _p(grad), _p(state_sum), param.numel(), float(lr), float(weight_decay),
                               float(eps), float(grad_scale), _p(grad_scale_dev), _stream())
    _lib.check(st, "lstc_adagrad_step")
    LAUNCHES.add(1)


# ------------------------------------------------------------------------------------------ multi-tensor launches
def _ptr_table(ptrs):
    import ctypes
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


def _i64_table(vals):
    import ctypes
    return (ctypes.c_int64 * len(vals))(*vals)


def multi_pack_bf16(srcs, flat: torch.Tensor, offsets) -> None:
    """fp32 tensors `srcs` -> bf16 `flat[offsets[i] : offsets[i] + srcs[i].numel()]` in one launch (offsets in elements,
    multiples of 8)."""
    lib = _lib.load()
    _cuda(flat, "flat", BF16)
    for t in srcs:
        _cuda(t, "src", F32)
        if not t.is_contiguous():
            raise RuntimeError("lstc_vad_b200.multi_pack_bf16: sources must be contiguous")
    n = len(srcs)
    base = flat.data_ptr()
    st = lib.lstc_multi_pack_bf16(_ptr_table([t.data_ptr() for t in srcs]), _i64_table([t.numel() for t in srcs]), n,
                                  _ptr_table([base + 2 * int(o) for o in offsets]), _stream())
    _lib.check(st, "lstc_multi_pack_bf16")
    LAUNCHES.add((n + 47) // 48)


def multi_unpack_bf16(flat: torch.Tensor, offsets, dsts) -> None:
    """bf16 `flat[offsets[i] : ...]` -> the fp32 tensors `dsts`, in one launch."""
    lib = _lib.load()
    _cuda(flat, "flat", BF16)
    for t in dsts:
        _cuda(t, "dst", F32)
        if not t.is_contiguous():
            raise RuntimeError("lstc_vad_b200.multi_unpack_bf16: destinations must be contiguous")
    n = len(dsts)
    base = flat.data_ptr()
    st = lib.lstc_multi_unpack_bf16(_ptr_table([base + 2 * int(o) for o in offsets]),
                                    _i64_table([t.numel() for t in dsts]), n,
                                    _ptr_table([t.data_ptr() for t in dsts]), _stream())
    _lib.check(st, "lstc_multi_unpack_bf16")
    LAUNCHES.add((n + 47) // 48)


def multi_adagrad(params, grads, states, lrs, weight_decay: float, eps: float = 1e-10, grad_scale: float = 1.0,
                  grad_scale_dev: Optional[torch.Tensor] = None) -> None:
    """torch.optim.Adagrad step of every (param, grad, state) triple in ONE launch per 48 tensors.  `grads` are fp32
    tensors, or all bf16 (slices of a reduced data-parallel bucket)."""
    import ctypes
    lib = _lib.load()
    n = len(params)
    if n == 0:
        return
    g_bf16 = grads[0].dtype == BF16
    for p_, g_, s_ in zip(params, grads, states):
        _cuda(p_, "param", F32)
        _cuda(s_, "state", F32)
        _cuda(g_, "grad", BF16 if g_bf16 else F32)
        if not (p_.is_contiguous() and g_.is_contiguous() and s_.is_contiguous()):
            raise RuntimeError("lstc_vad_b200.multi_adagrad: tensors must be contiguous")
        if g_.numel() != p_.numel() or s_.numel() != p_.numel():
            raise RuntimeError("lstc_vad_b200.multi_adagrad: size mismatch")
    st = lib.lstc_multi_adagrad(_ptr_table([t.data_ptr() for t in params]), _ptr_table([t.data_ptr() for t in grads]),
                                1 if g_bf16 else 0, _ptr_table([t.data_ptr() for t in states]),
                                _i64_table([t.numel() for t in params]), (ctypes.c_float * n)(*[float(x) for x in lrs]),
                                n, float(weight_decay), float(eps), float(grad_scale), _p(grad_scale_dev), _stream())
    _lib.check(st, "lstc_multi_adagrad")
    LAUNCHES.add((n + 47) // 48)
