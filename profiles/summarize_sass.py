#!/usr/bin/env python
"""Static evidence for the hot kernels: ptxas -v resource usage and counts of the SASS mnemonics that matter
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier ops, HMMA = mma.sync, LDSM = ldmatrix,
LDGSTS = cp.async).  Needs only nvcc / cuobjdump (no GPU):
    python profiles/summarize_sass.py > profiles/r2_ptxas_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gemm_tcgen05": ["gemm_bf16_tcgen05_2cta_kernelILi256ELb0ELb0", "gemm_bf16_tcgen05_2cta_kernelILi256ELb0ELb1",
                     "gemm_bf16_tcgen05_2cta_kernelILi256ELb1ELb1", "gemm_bf16_tcgen05_kernelILi256ELb0ELb0"],
    "attention_tc": ["attn_tc64_kernelILi64ELi256ELb0", "attn_tc64_kernelILi64ELi256ELb1", "attn_tc64_kernelILi32ELi256ELb0",
                     "attn_tc64_kernelILi32ELi256ELb1", "attn_tc128_kernelILi256ELb0", "attn_tc128_kernelILi256ELb1"],
    "attention": ["attn_fwd_kernelILi64ELi256ELi3", "attn_bwd_kernelILi64ELi256ELi3", "attn_fwd_kernelILi96ELi256ELi3",
                  "attn_bwd_kernelILi96ELi256ELi3", "attn_fwd_kernelILi32ELi256ELi3", "attn_bwd_kernelILi32ELi256ELi3"],
    "layernorm": ["ln_fwd_kernelILi1ELi4ELi6ELb0ELb0", "ln_fwd_kernelILi1ELi4ELi6ELb0ELb1",
                  "ln_bwd_kernelILi1ELi2ELi4ELb0ELb0ELb0", "ln_bwd_kernelILi1ELi2ELi4ELb1ELb0ELb0"],
}
KEYS = ["UTCHMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "SYNCS", "HMMA", "LDSM", "LDGSTS", "MUFU", "BAR",
        "RED", "SHFL", "STG", "LDG"]


def main():
    print("# static evidence for the hot kernels (nvcc, -gencode arch=compute_100a,code=sm_100a -O3; ptxas -v and cuobjdump -sass)")
    print("# kernel | registers | spill st/ld [B] | static smem [B] | SASS instruction count and mnemonics that matter")
    with tempfile.TemporaryDirectory() as tmp:
        for f, names in WANT.items():
            cubin = os.path.join(tmp, f + ".cubin")
            r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v",
                                "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
                                "-cubin", "-o", cubin, os.path.join(ROOT, "lstc_vad_b200", "csrc", f + ".cu")],
                               capture_output=True, text=True)
            info = {}
            for b in re.split(r"ptxas info\s+: Compiling entry function '", r.stderr)[1:]:
                nm = b.split("'")[0]
                m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
                rg = re.search(r"Used (\d+) registers", b)
                sm = re.search(r"(\d+) bytes smem", b)
                info[nm] = (rg.group(1) if rg else "?", m.group(1) if m else "?", m.group(2) if m else "?",
                            sm.group(1) if sm else "0")
            for n in names:
                full = [k for k in info if n in k]
                if not full:
                    continue
                k = full[0]
                sass = subprocess.run(["cuobjdump", "-sass", "-fun", k, cubin], capture_output=True, text=True).stdout
                c = collections.Counter()
                for line in sass.split("\n"):
                    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
                    if m:
                        c[m.group(2)] += 1
                mn = " ".join(f"{kk}={c[kk]}" for kk in KEYS if c[kk])
                print(f"{n} | {info[k][0]} | {info[k][1]}/{info[k][2]} | {info[k][3]} | total={sum(c.values())} {mn}")


if __name__ == "__main__":
    main()
