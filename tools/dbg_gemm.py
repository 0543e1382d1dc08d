import torch, math, sys
sys.path.insert(0, '.')  # run from the repo root
from lstc_vad_b200 import ops
from tests._util import report
torch.manual_seed(0)
def rnd(shape, s=1.0, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(shape, generator=g, device='cuda') * s).to(torch.bfloat16)
for (M, N, K) in [(304, 6144, 2048), (304, 2048, 2048), (304, 4096, 2048), (304, 2048, 4096), (300, 2048, 2048), (320, 2048, 2048), (384, 2048, 2048), (304, 256, 64), (304, 256, 512), (304, 512, 2048), (2432, 2048, 2048)]:
    a, b = rnd((M, K), 1, 1), rnd((N, K), 0.05, 2)
    ref = a.float() @ b.float().t()
    got = ops.gemm(a, b, out_dtype=torch.float32)
    ok, msg = report(f"kk {M}x{N}x{K}", got, ref, 1e-3, 1e-2)
    res = rnd((M, N), 1, 3); bias = torch.randn(N, device='cuda')
    got2 = ops.gemm(a, b, bias=bias, relu=True, residual=res)
    report(f"kk+epi {M}x{N}x{K}", got2, torch.relu(ref + bias) + res.float(), 2e-2, 5e-2)
    got3 = ops.gemm(a, b.t().contiguous(), b_mn=True, out_dtype=torch.float32)
    report(f"kmn {M}x{N}x{K}", got3, ref, 1e-3, 1e-2)
    if M % 8 == 0:
        got4 = ops.gemm(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True, out_dtype=torch.float32)
        report(f"mnmn {M}x{N}x{K}", got4, ref, 1e-3, 1e-2)
