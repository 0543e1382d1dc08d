"""Shared helpers for the module mirrors."""
import torch

from .. import functional as Fn
from ..ops import BF16, F32


def require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: lstc_vad_b200 runs on CUDA (sm_100a) only and has no CPU fallback; got a "
                           f"{x.device} tensor — move the module and its inputs to the GPU")


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == BF16:
        return x
    if x.dtype != F32:
        x = x.float()
    return Fn.ToBF16Fn.apply(x.contiguous())


def like_input(y: torch.Tensor, ref_dtype) -> torch.Tensor:
    """bf16 block output -> dtype of the caller's input (fp32 callers get fp32 back, as in the reference)."""
    if ref_dtype == BF16:
        return y
    return Fn.ToF32Fn.apply(y.contiguous())


def xavier_reset(module: torch.nn.Module) -> None:
    """`_reset_parameters` of the reference: xavier-uniform on every parameter with dim > 1, in registration
    order (models/Encoder.py:38-41, models/Classifier.py:15-18)."""
    for p in module.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
