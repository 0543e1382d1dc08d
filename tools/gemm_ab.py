"""Interleaved A/B timing of GEMM variants at the LTN-SHT layer shapes (one process, alternating A B A B ..., median of the
per-round time ratios): run-to-run clock / power-cap drift on a B200 is several per cent, larger than the effects compared.

    python tools/gemm_ab.py [--rounds 12] [--reps 8]

Pairs: bf16 outputs through TMA stores vs register stores (LSTC_GEMM_TMA_STORE flipped between launches), and input
gradients against the row-major weight (MN-major B operand) vs a transposed bf16 copy (K-major B operand).
"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lstc_vad_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--rows", type=int, default=1280 * 49)
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

    def env(v, fn):
        def run():
            os.environ["LSTC_GEMM_TMA_STORE"] = v
            return fn()
        return run

    fwd = {
        "fwd qkv bias": lambda: ops.gemm(x, wqkv, bias=bqkv),
        "fwd out-proj drop+res": lambda: ops.gemm(x, wo, residual=dy, dropout=drop),
        "fwd ffn w1 bias+relu": lambda: ops.gemm(x, w1, bias=b1, relu=True),
        "fwd ffn w2 bias+drop+res": lambda: ops.gemm(h, w2, bias=b2, residual=x, dropout=drop),
    }
    dg = {
        "dgrad ffn w2 relu-mask": (lambda: ops.gemm(dy, w2, b_mn=True, relu_mask=mask), lambda: ops.gemm(dy, w2t, relu_mask=mask)),
        "dgrad ffn w1": (lambda: ops.gemm(dh, w1, b_mn=True), lambda: ops.gemm(dh, w1t)),
        "dgrad out-proj": (lambda: ops.gemm(dy, wo, b_mn=True), lambda: ops.gemm(dy, wot)),
        "dgrad qkv": (lambda: ops.gemm(dqkv, wqkv, b_mn=True), lambda: ops.gemm(dqkv, wqkvt)),
    }
    pairs = []
    for n, f in fwd.items():
        pairs.append((n + ": register stores -> TMA stores", env("0", f), env("1", f)))
    for n, (fm, ft) in dg.items():
        pairs.append((n + ": MN-major B, register stores -> TMA stores", env("0", fm), env("2", fm)))
        pairs.append((n + ": K-major B (W^T), register stores -> TMA stores", env("0", ft), env("1", ft)))
        pairs.append((n + ": MN-major B (reg) -> K-major B (TMA)", env("0", fm), env("1", ft)))
        pairs.append((n + ": MN-major B (reg) -> K-major B (reg)", env("0", fm), env("0", ft)))
    print(f"# rows {R}, {args.rounds} rounds x {args.reps} launches per side; ratio = time(B) / time(A), < 1 means B is faster")
    for name, fa, fb in pairs:
        for _ in range(2):
            fa(); fb()
        torch.cuda.synchronize()
        ra, rb, ratios = [], [], []
        for r in range(args.rounds):
            if r % 2 == 0:
                ta = timed(fa, args.reps); tb = timed(fb, args.reps)
            else:
                tb = timed(fb, args.reps); ta = timed(fa, args.reps)
            ra.append(ta); rb.append(tb); ratios.append(tb / ta)
        print(f"{name:70s} A {statistics.median(ra):7.4f} ms  B {statistics.median(rb):7.4f} ms  "
              f"median ratio {statistics.median(ratios):6.3f}  (min {min(ratios):.3f} max {max(ratios):.3f})", flush=True)
    os.environ.pop("LSTC_GEMM_TMA_STORE", None)


if __name__ == "__main__":
    main()
