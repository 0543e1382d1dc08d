"""Pipeline accounting of the tcgen05 attention kernels: build the library with the wait counters compiled in

    LSTC_NVCC_EXTRA=-DLSTC_ATTN_TIMING python -m lstc_vad_b200.build --force

and run this script: CTA (0,0) of one backward and one forward launch prints, per role (TMA producers, MMA issuers,
softmax warps, epilogue groups), its total cycles and the cycles it spent waiting on each barrier class - which stage of
the pipeline is the critical one.  (Round 2 used it to find that a single MMA-issuing thread was busy 80 % of the time:
every product costs ~1 us of issue work for ~0.1 us of tensor time.)  LL=<tokens per window> selects the length.
Rebuild without the flag afterwards."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from lstc_vad_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
W, L, H, dk = 1280, int(os.environ.get("LL", "81")), 8, 256
rows = W * L
qkv = torch.randn(rows, 3 * H * dk, device=dev).to(torch.bfloat16)
do = torch.randn(rows, H * dk, device=dev).to(torch.bfloat16)
bias = torch.randn(H, L, L, device=dev) * 0.1
for i in range(2):
    ops.attn_bwd(qkv, do, W, L, H, dk, bias, 1 / 16.0, (0.2, 1, 0), True)
torch.cuda.synchronize()
print("=== BWD", flush=True)
ops.attn_bwd(qkv, do, W, L, H, dk, bias, 1 / 16.0, (0.2, 1, 0), True)
torch.cuda.synchronize()
print("=== FWD", flush=True)
ops.attn_fwd(qkv, W, L, H, dk, bias, 1 / 16.0, (0.2, 1, 0))
torch.cuda.synchronize()
