"""ctypes binding of the C-ABI declared in include/lstc_vad_b200.h.

The library is required: importing a kernel entry point when ``liblstc_vad_b200.so`` is missing raises —
there is no CPU or PyTorch fallback on the product path.
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "_C" / "liblstc_vad_b200.so"

_P = c_void_p
_I = c_int
_L = c_int64
_U = c_uint64
_F = c_float

# name -> (restype, argtypes).  Order and types mirror include/lstc_vad_b200.h exactly.
SIGNATURES: dict[str, tuple] = {
    "lstc_abi_version": (_I, []),
    "lstc_last_error": (c_char_p, []),
    "lstc_set_rng_step": (_I, [_P]),
    "lstc_gemm_bf16": (_I, [_P, _L, _I, _P, _L, _I, _L, _L, _L, _P, _L, _I, _P, _I, _P, _L, _P, _L, _F, _U, _U,
                            _I, _I, _P, _P]),
    "lstc_gemm_bf16_fuses_colsum": (_I, [_L, _L, _L, _I]),
    "lstc_attn_fwd": (_I, [_P, _L, _L, _I, _I, _I, _P, _F, _F, _U, _U, _P, _L, _P, _P]),
    "lstc_attn_bwd": (_I, [_P, _L, _P, _L, _L, _I, _I, _I, _P, _F, _F, _U, _U, _P, _L, _P, _P]),
    "lstc_attn_cls_fwd": (_I, [_P, _L, _P, _P, _L, _L, _I, _I, _I, _F, _F, _U, _U, _P, _L, _P]),
    "lstc_attn_cls_bwd": (_I, [_P, _L, _P, _P, _L, _P, _L, _L, _I, _I, _I, _F, _F, _U, _U, _P, _L, _P, _P, _L, _P]),
    "lstc_add_rows_bf16": (_I, [_P, _L, _P, _L, _L, _L, _P]),
    "lstc_relbias_gather": (_I, [_P, _P, _L, _I, _I, _L, _P, _P]),
    "lstc_relbias_scatter": (_I, [_P, _P, _L, _I, _I, _L, _P, _P]),
    "lstc_layernorm_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _L, _L, _F, _P]),
    "lstc_layernorm_bwd_workspace": (_L, [_L, _L]),
    "lstc_layernorm_bwd": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _F, _U, _U, _P, _P, _P, _P, _L, _L, _P]),
    "lstc_cls_prepend_fwd": (_I, [_P, _I, _P, _P, _F, _U, _U, _P, _L, _L, _L, _P]),
    "lstc_cls_prepend_bwd": (_I, [_P, _I, _F, _U, _U, _P, _P, _P, _L, _L, _L, _P]),
    "lstc_head_tail_fwd": (_I, [_P, _L, _I, _P, _P, _P, _P, _I, _I, _F, _U, _U, _P, _P, _P]),
    "lstc_head_tail_bwd": (_I, [_P, _P, _P, _P, _L, _I, _I, _F, _U, _U, _P, _P, _P, _P]),
    "lstc_mil_loss": (_I, [_P, _L, _I, _I, _I, _I, _F, _L, _P, _P, _P, _P]),
    "lstc_soft_ce_loss": (_I, [_P, _P, _L, _I, _P, _P, _P]),
    "lstc_bce_loss": (_I, [_P, _P, _L, _I, _F, _F, _P, _P, _P]),
    "lstc_threshold_labels": (_I, [_P, _F, _P, _L, _P]),
    "lstc_weighted_auc": (_I, [_P, _P, _P, _L, _P, _P]),
    "lstc_cast_f32_to_bf16": (_I, [_P, _P, _L, _P]),
    "lstc_cast_bf16_to_f32": (_I, [_P, _P, _L, _P]),
    "lstc_cast_f32_to_bf16_transposed": (_I, [_P, _L, _P, _L, _L, _L, _P]),
    "lstc_colsum_workspace": (_L, [_L, _L]),
    "lstc_colsum_bf16": (_I, [_P, _L, _L, _L, _P, _P, _P]),
    "lstc_dropout_apply_bf16": (_I, [_P, _P, _L, _L, _F, _U, _U, _P]),
    "lstc_dropout_mask": (_I, [_P, _L, _L, _F, _U, _U, _P]),
    "lstc_scale_by_device_scalar": (_I, [_P, _P, _P, _L, _P]),
    "lstc_gather_rows": (_I, [_P, _L, _P, _L, _P, _P]),
    "lstc_segment_mean": (_I, [_P, _P, _I, _L, _I, _I, _I, _P, _P]),
    "lstc_sumsq_accumulate": (_I, [_P, _L, _P, _P]),
    "lstc_clip_coef": (_I, [_P, _F, _P, _P]),
    "lstc_adagrad_step": (_I, [_P, _P, _P, _L, _F, _F, _F, _F, _P, _P]),
    "lstc_multi_pack_bf16": (_I, [_P, _P, _I, _P, _P]),
    "lstc_multi_unpack_bf16": (_I, [_P, _P, _I, _P, _P]),
    "lstc_multi_adagrad": (_I, [_P, _P, _I, _P, _P, _P, _I, _F, _F, _F, _P, _P]),
}

_lib = None
_lock = threading.Lock()


class LstcKernelError(RuntimeError):
    """A C-ABI entry point returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Loads (once) and returns the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -m lstc_vad_b200.build` "
                    "(lstc_vad_b200 has no CPU / PyTorch fallback)")
            lib = ctypes.CDLL(str(LIB_PATH))
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            if lib.lstc_abi_version() != 2:
                raise ImportError("liblstc_vad_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().lstc_last_error()
        raise LstcKernelError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")
