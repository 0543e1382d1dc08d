"""lstc_vad_b200 — B200 (sm_100a) implementation of the LSTC_VAD transformer hot path.

Public surface:
  lstc_vad_b200.models      drop-in mirror of the reference `models/` package
  lstc_vad_b200.losses      get_MIL_loss / get_CE_loss / get_BCE_loss / threshold_pseudo_labels
  lstc_vad_b200.ops         tensor-level wrappers of the C-ABI kernels (include/lstc_vad_b200.h)
  lstc_vad_b200.harness     batched train step, data-parallel sharding, sharded pseudo-label generation

Importing the package does not load the CUDA library; the first kernel call does, and raises if
`lstc_vad_b200/_C/liblstc_vad_b200.so` has not been built (`python -m lstc_vad_b200.build`).
"""
__version__ = "0.1.0"

from .functional import invalidate_weight_cache, set_dropout_stream  # noqa: F401
