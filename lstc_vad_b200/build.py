"""Builds the sm_100a C-ABI library ``lstc_vad_b200/_C/liblstc_vad_b200.so`` with nvcc.

No torch headers are involved: the library exposes a plain C ABI (include/lstc_vad_b200.h) and is loaded
with ctypes.  ``python -m lstc_vad_b200.build`` (or ``__graft_entry__.build()``) rebuilds when any source
is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OUT_DIR = PKG_DIR / "_C"
LIB_PATH = OUT_DIR / "liblstc_vad_b200.so"
INCLUDE = PKG_DIR.parent / "include"

SOURCES = ["runtime.cu", "gemm_tcgen05.cu", "attention.cu", "attention_tc.cu", "attention_cls.cu", "layernorm.cu", "elementwise.cu", "multitensor.cu", "heads.cu", "loss.cu"]

HEADERS = [CSRC / "common.cuh", CSRC / "ptx_sm100.cuh", CSRC / "attention_params.cuh", INCLUDE / "lstc_vad_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the lstc_vad_b200 CUDA library cannot be built")


def _deps() -> list[Path]:
    return [CSRC / s for s in SOURCES] + HEADERS


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(d.stat().st_mtime > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    common_t = max(h.stat().st_mtime for h in HEADERS)

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src.replace(".cu", ".o"))
        s = CSRC / src
        if not force and obj.exists() and obj.stat().st_mtime > max(s.stat().st_mtime, common_t):
            return obj
        extra = os.environ.get("LSTC_NVCC_EXTRA", "").split()  # e.g. -DLSTC_ATTN_TIMING for the pipeline accounting
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-c", str(s), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(f"built {p}")
