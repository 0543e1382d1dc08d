from torch import nn

from .. import functional as Fn
from ._common import like_input, require_cuda, to_bf16


class PositionwiseFeedForward(nn.Module):
    """w_2(relu(w_1 x)) + dropout + residual (+ LayerNorm eps 1e-6) — reference models/FFN.py:6-22.

    Two tcgen05 GEMMs: bias+ReLU fused into the first epilogue, bias+dropout+residual into the second;
    the [tokens, d_hid] intermediate is bf16."""

    def __init__(self, d_in, d_hid, dropout=0.1, layerNorm=True):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.layer_norm = nn.LayerNorm(d_in, eps=1e-6)
        self.dropout = nn.Dropout(dropout)
        self.layerNorm_flag = layerNorm

    def _forward_bf16(self, x, out_f32=False):
        """x bf16 -> bf16, or fp32 when `out_f32` and the block ends in its LayerNorm (which then writes fp32)."""
        cfg = Fn.FFNConfig(layer_norm=self.layerNorm_flag == True,  # noqa: E712 (reference compares with ==)
                           drop=Fn.next_dropout(self.dropout.p, self.training),
                           out_f32=bool(out_f32) and self.layerNorm_flag == True)  # noqa: E712
        return Fn.FFNBlockFn.apply(x, self.w_1.weight, self.w_1.bias, self.w_2.weight, self.w_2.bias,
                                   self.layer_norm.weight, self.layer_norm.bias, cfg)

    def forward(self, x):
        require_cuda(x, "PositionwiseFeedForward")
        return like_input(self._forward_bf16(to_bf16(x)), x.dtype)
