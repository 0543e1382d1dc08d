from torch import nn

from ._common import like_input, require_cuda, to_bf16
from .FFN import PositionwiseFeedForward
from .MultiHeadAttention import MultiHeadAttention


class EncoderLayer(nn.Module):
    """Self-attention block followed by the position-wise FFN — reference models/EncoderLayer.py:4-30."""

    def __init__(self, d_model, d_inner, n_head, d_k, d_v, MHA_attn_dropout=0.1, MHA_fc_dropout=0.1,
                 MHA_layerNorm=False, FFN_dropout=0.1, FFN_layerNorm=True, return_attn=False,
                 relative_pe=False, window_size=4, window_depth=3, conv_patch=False,
                 relative_pe_2D=False, FFN_need=True):
        super(EncoderLayer, self).__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v,
                                           attn_dropout=MHA_attn_dropout, fc_dropout=MHA_fc_dropout,
                                           layerNorm=MHA_layerNorm, relative_pe=relative_pe, window_size=window_size,
                                           window_depth=window_depth, conv_patch=conv_patch,
                                           relative_pe_2D=relative_pe_2D)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=FFN_dropout, layerNorm=FFN_layerNorm)
        self.FFN_need = FFN_need

    def _forward_bf16(self, x, return_attn=False, return_attn_v=False, out_f32=False):
        out, attn, v = self.slf_attn._forward_bf16(x, return_attn, return_attn_v)
        if self.FFN_need == True:  # noqa: E712
            out = self.pos_ffn._forward_bf16(out, out_f32=out_f32)
        return out, attn, v

    def _forward_cls_bf16(self, x):
        """x bf16 [W,L,D] -> bf16 [W,D]: CLS row of this layer's output (valid only for the LAST layer)."""
        out = self.slf_attn._forward_cls_bf16(x)
        if self.FFN_need == True:  # noqa: E712
            out = self.pos_ffn._forward_bf16(out.view(out.shape[0], 1, out.shape[1])).view(out.shape)
        return out

    def forward(self, enc_input, slf_attn_mask=None, return_attn=False, return_attn_v=False):
        require_cuda(enc_input, "EncoderLayer")
        if slf_attn_mask is not None:
            raise NotImplementedError("attention masks are not supported (no reference caller passes one)")
        out, attn, v = self._forward_bf16(to_bf16(enc_input), return_attn, return_attn_v)
        out = like_input(out, enc_input.dtype)
        if return_attn_v == True:  # noqa: E712
            return out, attn, v
        return out, (attn if return_attn else None)
