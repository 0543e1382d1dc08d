"""Times the GEMMs of one LTN-SHT EncoderLayer alone (forward, dgrad, wgrad; the epilogues the train step uses).

    python tools/gemm_bench.py [--reps 10] [--only wgrad] [--rows 62720] [--check]

Per launch: mean ms (CUDA events around `reps` back-to-back launches after 2 warm-ups), TFLOP/s and the fraction of
MEASURED_PEAKS.json's sustained bf16 rate.  `--check` compares every result with a torch fp32 product (rel-L2).
Run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` with `--reps 1` to get the DRAM traffic per
launch (env knobs: LSTC_GEMM_TMA_STORE=0 register stores instead of TMA stores, LSTC_GEMM_L2_HINTS=0 plain operand
loads and N-first tile order).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lstc_vad_b200 import functional as Fn  # noqa: E402
from lstc_vad_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def peak_tflops():
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained") or d.get("bf16_tflops") or 1409.8)
    except Exception:
        return 1409.8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=1280 * 49)
    ap.add_argument("--only", default="")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    R, D, Dh = args.rows, 2048, 4096
    g = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(BF16)
    x, h, dy, dh, dqkv = g(R, D), g(R, Dh), g(R, D), g(R, Dh), g(R, 3 * D)
    mask = torch.relu(g(R, Dh))
    wqkv, wo, w1, w2 = g(3 * D, D), g(D, D), g(Dh, D), g(D, Dh)
    wqkvt, wot, w1t, w2t = (w.t().contiguous() for w in (wqkv, wo, w1, w2))
    bqkv, b1, b2 = torch.randn(3 * D, device=dev), torch.randn(Dh, device=dev), torch.randn(D, device=dev)
    drop = (0.2, 1, 0)
    cases = [
        # name, fn, (M, N, K), torch reference (or None)
        ("fwd qkv      bias", lambda: ops.gemm(x, wqkv, bias=bqkv), (R, 3 * D, D), lambda: x.float() @ wqkv.float().t() + bqkv),
        ("fwd out-proj drop+res", lambda: ops.gemm(x, wo, residual=dy, dropout=drop), (R, D, D), None),
        ("fwd out-proj res", lambda: ops.gemm(x, wo, residual=dy), (R, D, D), lambda: x.float() @ wo.float().t() + dy.float()),
        ("fwd ffn w1   bias+relu", lambda: ops.gemm(x, w1, bias=b1, relu=True), (R, Dh, D),
         lambda: torch.relu(x.float() @ w1.float().t() + b1)),
        ("fwd ffn w2   bias+drop+res", lambda: ops.gemm(h, w2, bias=b2, residual=x, dropout=drop), (R, D, Dh), None),
        ("dgrad ffn w2 relu-mask", lambda: ops.gemm(dy, w2, b_mn=True, relu_mask=mask), (R, Dh, D),
         lambda: (dy.float() @ w2.float()) * (mask.float() > 0)),
        ("dgrad ffn w1", lambda: ops.gemm(dh, w1, b_mn=True), (R, D, Dh), lambda: dh.float() @ w1.float()),
        ("dgrad out-proj", lambda: ops.gemm(dy, wo, b_mn=True), (R, D, D), lambda: dy.float() @ wo.float()),
        ("dgrad qkv", lambda: ops.gemm(dqkv, wqkv, b_mn=True), (R, D, 3 * D), lambda: dqkv.float() @ wqkv.float()),
        # the same input gradients against transposed weight copies (K-major B, the forward's operand layout)
        ("dgradT ffn w2 relu-mask", lambda: ops.gemm(dy, w2t, relu_mask=mask), (R, Dh, D),
         lambda: (dy.float() @ w2.float()) * (mask.float() > 0)),
        ("dgradT ffn w1", lambda: ops.gemm(dh, w1t), (R, D, Dh), lambda: dh.float() @ w1.float()),
        ("dgradT out-proj", lambda: ops.gemm(dy, wot), (R, D, D), lambda: dy.float() @ wo.float()),
        ("dgradT qkv", lambda: ops.gemm(dqkv, wqkvt), (R, D, 3 * D), lambda: dqkv.float() @ wqkv.float()),
        ("wgrad ffn w2", lambda: Fn._wgrad(dy, h), (D, Dh, R), lambda: dy.float().t() @ h.float()),
        ("wgrad ffn w1", lambda: Fn._wgrad(dh, x), (Dh, D, R), lambda: dh.float().t() @ x.float()),
        ("wgrad out-proj", lambda: Fn._wgrad(dy, x), (D, D, R), lambda: dy.float().t() @ x.float()),
        ("wgrad qkv", lambda: Fn._wgrad(dqkv, x), (3 * D, D, R), lambda: dqkv.float().t() @ x.float()),
    ]
    peak = peak_tflops()
    knobs = {k: os.environ[k] for k in ("LSTC_GEMM_TMA_STORE", "LSTC_GEMM_L2_HINTS") if k in os.environ}
    print(f"# rows {R}, reps {args.reps}, knobs {knobs}", flush=True)
    tot_ms = tot_fl = 0.0
    for name, fn, (M, N, K), ref in cases:
        if args.only and args.only not in name:
            continue
        if args.check:
            got = fn().float()
            torch.cuda.synchronize()
            if ref is not None:
                want = ref()
                err = ((got - want).norm() / want.norm()).item()
                print(f"{name:30s} rel-L2 vs torch fp32 {err:.2e} {'OK' if err < 8e-3 else 'MISMATCH'}", flush=True)
                del want
            del got
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        fl = 2.0 * M * N * K
        tot_ms += ms
        tot_fl += fl
        print(f"{name:30s} {ms:8.4f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  {fl / ms / 1e9 / peak:5.2f} of sustained peak", flush=True)
    if tot_ms > 0:
        print(f"{'sum':30s} {tot_ms:8.4f} ms  {tot_fl / tot_ms / 1e9:8.1f} TFLOP/s  {tot_fl / tot_ms / 1e9 / peak:5.2f}", flush=True)


if __name__ == "__main__":
    main()
