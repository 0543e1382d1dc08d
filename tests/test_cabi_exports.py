"""CPU-only: the C-ABI library loads, exports every symbol include/lstc_vad_b200.h declares, and the ctypes
signatures in lstc_vad_b200/_lib.py agree with the header (argument counts and pointer/scalar kinds).
No kernel is launched here."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "lstc_vad_b200.h").read_text()


def header_decls():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    decls = {}
    for m in re.finditer(r"^\s*(const char\*|int64_t|int)\s+(lstc_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S | re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        decls[name] = (ret, params)
    return decls


def kind_of(param: str) -> str:
    if "*" in param:
        return "ptr"
    ty = param.rsplit(" ", 1)[0].strip()
    return {"int": "int", "int64_t": "i64", "uint64_t": "u64", "float": "f32"}[ty]


def test_header_declares_the_expected_entry_points():
    d = header_decls()
    for name in ("lstc_gemm_bf16", "lstc_attn_fwd", "lstc_attn_bwd", "lstc_layernorm_fwd", "lstc_layernorm_bwd",
                 "lstc_cls_prepend_fwd", "lstc_cls_prepend_bwd", "lstc_head_tail_fwd", "lstc_head_tail_bwd",
                 "lstc_mil_loss", "lstc_soft_ce_loss", "lstc_bce_loss", "lstc_threshold_labels", "lstc_adagrad_step",
                 "lstc_relbias_gather", "lstc_relbias_scatter", "lstc_last_error", "lstc_abi_version"):
        assert name in d, name
    assert len(d) >= 26


def test_library_exports_every_declared_symbol_and_signatures_match():
    import ctypes
    from lstc_vad_b200 import _lib, build
    build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    lib = _lib.load()
    assert lib.lstc_abi_version() == 2
    decls = header_decls()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    cmap = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "int", ctypes.c_int64: "i64",
            ctypes.c_uint64: "u64", ctypes.c_float: "f32"}
    for name, (ret, params) in decls.items():
        assert hasattr(lib, name), f"{name} not exported"
        res, argtypes = _lib.SIGNATURES[name]
        assert len(argtypes) == len(params), f"{name}: {len(argtypes)} ctypes args vs {len(params)} in the header"
        for i, (a, p) in enumerate(zip(argtypes, params)):
            assert cmap[a] == kind_of(p), f"{name} arg {i} ({p}): ctypes {cmap[a]}"
        assert {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char*": ctypes.c_char_p}[ret] == res


def test_argument_errors_are_reported_without_a_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from lstc_vad_b200 import _lib
    lib = _lib.load()
    st = lib.lstc_gemm_bf16(None, 8, 0, None, 8, 0, 1, 1, 1, None, 8, 0, None, 0, None, 0, None, 0, 0.0, 0, 0, 1, 0, None, None)
    assert st == 1 and b"null operand" in lib.lstc_last_error()
    st = lib.lstc_layernorm_fwd(1, 0, 1, 1, 1, 0, 1, 1, 4, 12, 1e-6, None)
    assert st == 1 and b"multiple of 8" in lib.lstc_last_error()
    with pytest.raises(_lib.LstcKernelError):
        _lib.check(st, "lstc_layernorm_fwd")


def test_gemm_colsum_capability_query_is_pure_host_logic(monkeypatch):
    """lstc_gemm_bf16_fuses_colsum: bf16 C, 16-byte pitch, at least one [32 x 64] store box - and the same answer decides
    whether lstc_gemm_bf16 accepts a colsum pointer (checked before any CUDA call)."""
    from lstc_vad_b200 import _lib
    lib = _lib.load()
    monkeypatch.delenv("LSTC_GEMM_TMA_STORE", raising=False)
    assert lib.lstc_gemm_bf16_fuses_colsum(62720, 4096, 4096, 0) == 1
    assert lib.lstc_gemm_bf16_fuses_colsum(62720, 4096, 4096, 1) == 0     # fp32 C
    assert lib.lstc_gemm_bf16_fuses_colsum(62720, 3027, 3027, 0) == 0     # pitch not a multiple of 8 elements
    assert lib.lstc_gemm_bf16_fuses_colsum(62720, 32, 32, 0) == 0         # narrower than one store box
    assert lib.lstc_gemm_bf16_fuses_colsum(16, 4096, 4096, 0) == 0        # shorter than one store box
    # a colsum pointer with an output the epilogue cannot fuse it for is an argument error, not a silent skip
    st = lib.lstc_gemm_bf16(16, 8, 0, 16, 8, 0, 64, 32, 64, 16, 32, 0, None, 0, None, 0, None, 0, 0.0, 0, 0, 1, 0, 16, None)
    assert st == 1 and b"fused column sums" in lib.lstc_last_error()
    monkeypatch.setenv("LSTC_GEMM_TMA_STORE", "0")                        # register-store epilogue: no staged tile to sum
    assert lib.lstc_gemm_bf16_fuses_colsum(62720, 4096, 4096, 0) == 0


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "lstc_vad_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in src.replace("no oracle", ""), f"{py} references the oracle"
