import torch
from torch import nn

from .. import functional as Fn
from ..ops import BF16
from ._common import like_input, require_cuda, xavier_reset
from .EncoderLayer import EncoderLayer


class Encoder(nn.Module):
    """Spatio / temporal transformer encoder over windows of pre-extracted I3D tokens — reference
    models/Encoder.py:4-74.  Same constructor, forward signature, parameter names / shapes / registration
    order and state_dict layout; the math runs in the lstc_vad_b200 CUDA kernels (bf16 activations, fp32
    accumulation and statistics) and the result is returned in the caller's dtype."""

    def __init__(self, n_layers, n_head, d_k, d_v, d_model, d_inner,
                 MHA_attn_dropout=0.1, MHA_fc_dropout=0.1, MHA_layerNorm=False,
                 FFN_dropout=0.1, FFN_layerNorm=True,
                 weight_init=True, CLS_learned=False, position_dropout=0.1, position_encoding=False,
                 max_position_tokens=100,
                 relative_pe=False, window_size=4, window_depth=3, conv_patch=False, input_layerNorm=False,
                 relative_pe_2D=False,
                 FFN_need=True):
        super().__init__()
        self.CLS_learned = CLS_learned
        if CLS_learned == True:  # noqa: E712
            self.cls_token = nn.Parameter(torch.randn(1, 1, d_model))
        self.position_encoding = position_encoding
        if position_encoding == True:  # noqa: E712
            self.position_dropout = nn.Dropout(position_dropout)
            self.position_enc = nn.Parameter(torch.randn(1, max_position_tokens, d_model))

        self.layer_stack = nn.ModuleList([
            EncoderLayer(d_model, d_inner, n_head, d_k, d_v,
                         MHA_attn_dropout=MHA_attn_dropout, MHA_fc_dropout=MHA_fc_dropout, MHA_layerNorm=MHA_layerNorm,
                         FFN_dropout=FFN_dropout, FFN_layerNorm=FFN_layerNorm,
                         relative_pe=relative_pe, window_size=window_size,
                         window_depth=window_depth, conv_patch=conv_patch,
                         relative_pe_2D=relative_pe_2D, FFN_need=FFN_need)
            for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.input_layerNorm = input_layerNorm
        if weight_init == True:  # noqa: E712
            self._reset_parameters()

    def _reset_parameters(self):
        xavier_reset(self)

    def _embed(self, enc_output):
        in_dtype = enc_output.dtype
        x = enc_output if in_dtype in (torch.float32, BF16) else enc_output.float()
        if self.input_layerNorm == True:  # noqa: E712
            x = Fn.LayerNormFn.apply(x, self.layer_norm.weight, self.layer_norm.bias, True)
        pos = None
        drop = Fn.NO_DROPOUT
        if self.position_encoding == True:  # noqa: E712
            pos = self.position_enc
            drop = Fn.next_dropout(self.position_dropout.p, self.training)
        return Fn.ClsPrependFn.apply(x, self.cls_token if self.CLS_learned == True else None, pos, drop)  # noqa: E712

    def forward_cls(self, enc_output):
        """Opt-in fast path (NOT part of the reference API): returns `self.forward(enc_output)[:, 0, :]` — the only
        part of the output any reference caller consumes — without computing the last layer's out-projection / FFN /
        LayerNorms and Q projection on the other L-1 tokens (exact dead-work elimination, ~25 % of the FLOPs at L=49).
        [W, L0, D] -> [W, D] in the caller's dtype."""
        require_cuda(enc_output, "Encoder")
        if enc_output.dim() != 3:
            raise RuntimeError(f"Encoder expects [windows, tokens, d_model], got {tuple(enc_output.shape)}")
        x = self._embed(enc_output)
        n = len(self.layer_stack)
        for i, enc_layer in enumerate(self.layer_stack):
            if i < n - 1:
                x, _, _ = enc_layer._forward_bf16(x)
            else:
                x = enc_layer._forward_cls_bf16(x)
        return like_input(x, BF16 if enc_output.dtype == BF16 else torch.float32)

    def forward(self, enc_output, src_mask=None, return_attn=False, return_attn_v=False):
        require_cuda(enc_output, "Encoder")
        if src_mask is not None:
            raise NotImplementedError("attention masks are not supported (no reference caller passes one)")
        if enc_output.dim() != 3:
            raise RuntimeError(f"Encoder expects [windows, tokens, d_model], got {tuple(enc_output.shape)}")
        in_dtype = enc_output.dtype
        x = self._embed(enc_output)

        attn_list, v_list = [], []
        want_f32 = in_dtype != BF16
        n_layers = len(self.layer_stack)
        for li, enc_layer in enumerate(self.layer_stack):
            # the last block's LayerNorm writes the caller's fp32 output directly (no separate cast pass)
            x, attn, v = enc_layer._forward_bf16(x, return_attn, return_attn_v,
                                                 out_f32=want_f32 and li == n_layers - 1)
            if return_attn or return_attn_v:
                attn_list.append(attn)
            if return_attn_v:
                v_list.append(v)
        out = x if (want_f32 and x.dtype == torch.float32) else like_input(x, BF16 if in_dtype == BF16 else torch.float32)
        if return_attn_v == True:  # noqa: E712
            return out, attn_list, v_list
        if return_attn:
            return out, attn_list
        return out
