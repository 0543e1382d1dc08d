import sys, os, torch
sys.path.insert(0, os.getcwd())
from lstc_vad_b200 import ops
from tools.kernel_bench import timeit
dev = torch.device("cuda:0"); torch.manual_seed(0)
W, L, H, dk = 1280, int(os.environ.get("LL", "49")), 8, 256
rows = W * L
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
qkv = torch.randn(rows, 3 * H * dk, device=dev).to(torch.bfloat16)
do = torch.randn(rows, H * dk, device=dev).to(torch.bfloat16)
bias = torch.randn(H, L, L, device=dev) * 0.1
scale = 1 / 16.0
ab = rows * H * dk * 2
for name, fn, nb in [
    ("bwd plain", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, None, scale), 7 * ab),
    ("bwd bias", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, bias, scale), 7 * ab),
    ("bwd bias+dbias", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, bias, scale, ops.NO_DROPOUT, True), 7 * ab),
    ("bwd dropout", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, None, scale, (0.2, 1, 0)), 7 * ab),
    ("bwd all", lambda: ops.attn_bwd(qkv, do, W, L, H, dk, bias, scale, (0.2, 1, 0), True), 7 * ab),
    ("fwd plain", lambda: ops.attn_fwd(qkv, W, L, H, dk, None, scale), 4 * ab),
    ("fwd all", lambda: ops.attn_fwd(qkv, W, L, H, dk, bias, scale, (0.2, 1, 0)), 4 * ab),
]:
    ms = timeit(fn, 20, flush)
    print(f"L={L} {name:18s} {ms:.4f} ms {nb/ms/1e6:8.1f} GB/s {nb/ms/1e6/6544:.2f}", flush=True)
