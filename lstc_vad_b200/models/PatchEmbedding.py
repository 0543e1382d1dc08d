import torch
from torch import nn

from .. import functional as Fn
from ._common import like_input, require_cuda


class PatchEmbedding(nn.Module):
    """CLS-token prepend (token mean or learned token) — reference models/PatchEmbedding.py:4-19.
    The reference never instantiates it (Encoder inlines the same logic); kept for API parity."""

    def __init__(self, embed_dim=2048, CLS_learned=False):
        super().__init__()
        self.CLS_learned = CLS_learned
        if CLS_learned == True:  # noqa: E712
            self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))

    def forward(self, feats):
        require_cuda(feats, "PatchEmbedding")
        x = feats if feats.dtype in (torch.float32, torch.bfloat16) else feats.float()
        out = Fn.ClsPrependFn.apply(x, self.cls_token if self.CLS_learned == True else None, None,  # noqa: E712
                                    Fn.NO_DROPOUT)
        return like_input(out, x.dtype)
