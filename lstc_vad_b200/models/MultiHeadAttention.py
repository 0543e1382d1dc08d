import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from .. import functional as Fn
from ._common import like_input, require_cuda, to_bf16


class ScaledDotProductAttention(nn.Module):
    """Kept importable for API parity (reference models/MultiHeadAttention.py:9-23); the reference never
    instantiates it — MultiHeadAttention inlines its own attention — and neither does this package."""

    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, q, k, v, mask=None, relative_pe=False, window_size=4):
        raise NotImplementedError("ScaledDotProductAttention is dead code in the reference; use MultiHeadAttention")


def _rel_index_3d(wd: int, ws: int) -> torch.Tensor:
    """int64 [(wd*ws*ws)^2 as (n,n)] pair-wise relative position index, tokens ordered depth, row, column
    (reference models/MultiHeadAttention.py:59-73)."""
    d, h, w = np.meshgrid(np.arange(wd), np.arange(ws), np.arange(ws), indexing="ij")
    d, h, w = d.ravel(), h.ravel(), w.ravel()
    s = 2 * ws - 1
    idx = (d[:, None] - d[None, :] + wd - 1) * s * s + (h[:, None] - h[None, :] + ws - 1) * s \
        + (w[:, None] - w[None, :] + ws - 1)
    return torch.from_numpy(idx.astype(np.int64))


def _rel_index_2d(ws: int) -> torch.Tensor:
    """reference models/MultiHeadAttention.py:76-89."""
    h, w = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    h, w = h.ravel(), w.ravel()
    s = 2 * ws - 1
    idx = (h[:, None] - h[None, :] + ws - 1) * s + (w[:, None] - w[None, :] + ws - 1)
    return torch.from_numpy(idx.astype(np.int64))


class MultiHeadAttention(nn.Module):
    """Multi-head self-attention with Swin-style relative-position bias — reference
    models/MultiHeadAttention.py:25-132.  One fused-QKV tcgen05 GEMM, one fused attention kernel per
    (window, head), one out-projection GEMM with dropout + residual in its epilogue, one LayerNorm kernel."""

    def __init__(self, n_head, d_model, d_k, d_v, layerNorm=False,
                 attn_dropout=0.1, fc_dropout=0.1, relative_pe=False, window_size=3,
                 window_depth=3, conv_patch=False, relative_pe_2D=False):
        super().__init__()
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.layerNorm_flag = layerNorm
        self.d_model = d_model

        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)

        self.dropout = nn.Dropout(fc_dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)

        self.temperature = d_k ** 0.5
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.relative_pe_2D = relative_pe_2D
        self.relative_pe = relative_pe
        self.window_size = window_size
        self.window_depth = window_depth
        if relative_pe == True:  # noqa: E712
            self.relative_position_bias_table = nn.Parameter(
                torch.zeros((2 * window_depth - 1) * (2 * window_size - 1) * (2 * window_size - 1), n_head))
            self.register_buffer("relative_position_index", _rel_index_3d(window_depth, window_size))
            nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        if relative_pe_2D == True:  # noqa: E712
            self.relative_position_bias_table = nn.Parameter(
                torch.zeros((2 * window_size - 1) * (2 * window_size - 1), n_head))
            self.register_buffer("relative_position_index", _rel_index_2d(window_size))
            nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)

    def _forward_bf16(self, x, return_attn=False, return_attn_v=False):
        """x bf16 [W,L,D] -> (out bf16, attn fp32 | None, v fp32 [W,H,L,dv] | None)."""
        if self.d_k != self.d_v:
            raise NotImplementedError("lstc_vad_b200 attention requires d_k == d_v (all reference configs use 256)")
        rel_mode = 2 if self.relative_pe_2D == True else (1 if self.relative_pe == True else 0)  # noqa: E712
        cfg = Fn.MHAConfig(n_head=self.n_head, d_k=self.d_k, layer_norm=self.layerNorm_flag == True,  # noqa: E712
                           rel_mode=rel_mode,
                           attn_drop=Fn.next_dropout(self.attn_dropout.p, self.training),
                           fc_drop=Fn.next_dropout(self.dropout.p, self.training),
                           want_attn=bool(return_attn or return_attn_v))
        table = self.relative_position_bias_table if rel_mode else None
        index = self.relative_position_index if rel_mode else None
        out, attn = Fn.MHABlockFn.apply(x, self.w_qs.weight, self.w_ks.weight, self.w_vs.weight, self.fc.weight,
                                        self.layer_norm.weight, self.layer_norm.bias, table, index, cfg)
        v = None
        if return_attn_v:
            # diagnostic output only (reference returns the projected values, models/MultiHeadAttention.py:127-128)
            W, L, _ = x.shape
            with torch.no_grad():
                v = F.linear(x.float(), self.w_vs.weight).view(W, L, self.n_head, self.d_v).transpose(1, 2)
        return out, attn, v

    def _forward_cls_bf16(self, x):
        """Last-layer fast path: x bf16 [W,L,D] -> bf16 [W,D] (the CLS row of the block output)."""
        if self.d_k != self.d_v:
            raise NotImplementedError("lstc_vad_b200 attention requires d_k == d_v (all reference configs use 256)")
        cfg = Fn.MHAConfig(n_head=self.n_head, d_k=self.d_k, layer_norm=self.layerNorm_flag == True,  # noqa: E712
                           rel_mode=0,  # the CLS row / column of the rel-pos bias is zero
                           attn_drop=Fn.next_dropout(self.attn_dropout.p, self.training),
                           fc_drop=Fn.next_dropout(self.dropout.p, self.training))
        has_table = self.relative_pe == True or self.relative_pe_2D == True  # noqa: E712
        return Fn.MHAClsFn.apply(x, self.w_qs.weight, self.w_ks.weight, self.w_vs.weight, self.fc.weight,
                                 self.layer_norm.weight, self.layer_norm.bias,
                                 self.relative_position_bias_table if has_table else None, cfg)

    def forward(self, q, k, v, mask=None, return_attn=False, return_attn_v=False):
        require_cuda(q, "MultiHeadAttention")
        if mask is not None:
            raise NotImplementedError("attention masks are not supported (no reference caller passes one)")
        if not (k is q and v is q):
            raise NotImplementedError("lstc_vad_b200 implements self-attention only (q is k is v), the only form "
                                      "the reference uses (models/EncoderLayer.py:20-24)")
        out, attn, vv = self._forward_bf16(to_bf16(q), return_attn, return_attn_v)
        out = like_input(out, q.dtype)
        if return_attn_v == True:  # noqa: E712
            return out, attn, vv
        if return_attn == False:  # noqa: E712
            return out, None
        return out, attn
