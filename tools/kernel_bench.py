"""Times the non-GEMM kernels of one LTN-SHT layer alone (CUDA events, L2 flushed between repetitions).

    python tools/kernel_bench.py [--reps 20]

Prints per kernel: mean ms, algorithmic GB moved, GB/s and the fraction of MEASURED_PEAKS.json's HBM figure.
The sizes are the headline workload's: 1280 windows x 49 tokens x d_model 2048, 8 heads x 256.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstc_vad_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def hbm_peak():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbps", "hbm_GBps", "hbm_gb_s", "hbm"):
            if k in d:
                return float(d[k])
        for k, v in d.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v)
    except Exception:
        pass
    return 6544.0


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--L", type=int, default=49, help="tokens per window incl. CLS (49 SHT, 19 UCF, 81 UBnormal, 17 STN)")
    ap.add_argument("--W", type=int, default=1280)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    W, L, H, dk, D = args.W, args.L, 8, 256, 2048
    rows = W * L
    peak = hbm_peak()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    x = torch.randn(rows, D, device=dev).to(BF16)
    dy = torch.randn(rows, D, device=dev).to(BF16)
    g = torch.rand(D, device=dev) + 0.5
    b = torch.randn(D, device=dev)
    _, mean, rstd = ops.layernorm_fwd(x, g, b)
    dy32 = dy.float()
    bf = rows * D * 2
    cases = [
        ("ln_fwd bf16->bf16", lambda: ops.layernorm_fwd(x, g, b), 2 * bf),
        ("ln_fwd bf16->f32", lambda: ops.layernorm_fwd(x, g, b, 1e-6, F32), 3 * bf),
        ("ln_bwd plain", lambda: ops.layernorm_bwd(dy, x, g, mean, rstd), 3 * bf),
        ("ln_bwd dropout", lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dropout=(0.2, 1, 0)), 4 * bf),
        ("ln_bwd dropout+dxsum", lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dropout=(0.1, 1, 0), want_dxsum=True), 4 * bf),
        ("ln_bwd dy f32 dropout+dxsum", lambda: ops.layernorm_bwd(dy32, x, g, mean, rstd, dropout=(0.1, 1, 0), want_dxsum=True), 5 * bf),
    ]
    qkv = torch.randn(rows, 3 * H * dk, device=dev).to(BF16)
    do = torch.randn(rows, H * dk, device=dev).to(BF16)
    bias = torch.randn(H, L, L, device=dev) * 0.1
    scale = 1.0 / 16.0
    ab = rows * H * dk * 2
    cases += [
        ("attn_fwd dropout+bias", lambda: ops.attn_fwd(qkv, W, L, H, dk, bias, scale, (0.2, 1, 0)), 4 * ab),
        ("attn_fwd plain", lambda: ops.attn_fwd(qkv, W, L, H, dk, None, scale), 4 * ab),
        ("attn_bwd dropout+bias+dbias", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, bias, scale, (0.2, 1, 0), True), 7 * ab),
        ("attn_bwd plain", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, None, scale), 7 * ab),
    ]
    xf = torch.randn(W, L - 1, D, device=dev)
    cases += [
        ("cls_prepend_fwd f32->bf16", lambda: ops.cls_prepend_fwd(xf), W * (L - 1) * D * 4 + bf),
        ("colsum bf16", lambda: ops.colsum(x), bf),
    ]
    for name, fn, nbytes in cases:
        if args.only and args.only not in name:
            continue
        ms = timeit(fn, args.reps, flush)
        gbs = nbytes / ms / 1e6
        print(f"{name:32s} {ms:8.4f} ms  {nbytes / 1e9:6.3f} GB  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of HBM peak", flush=True)


if __name__ == "__main__":
    main()
