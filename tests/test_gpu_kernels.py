"""Per-kernel parity tests (GPU): every C-ABI kernel against the plain fp32 torch expression of the same op.

Tolerances: operands are bf16 (8 mantissa bits), accumulation and statistics are fp32.  A bf16-rounded output
carries a relative error of 2^-9 ~ 2e-3 per element; tests use rtol 1e-2..2e-2 on top of a small atol
scaled to the magnitude of the result.  Integer outputs (top-k indices, masks) are bit-exact.
"""
import math

import pytest
import torch

from tests._util import assert_close

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    from lstc_vad_b200 import ops as _ops
    return _ops


def _rand(shape, scale=1.0, seed=0, dtype=BF16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=F32) * scale).to(dtype)


# ------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (128, 256, 64),      # one tile, one k-block
    (128, 256, 256),     # one tile, pipeline wraps once
    (256, 512, 512),     # 2x2 tiles
    (300, 520, 200),     # ragged M, N, K
    (4900, 2048, 2048),  # LTN projection shape, 100 windows of 49 tokens
    (64, 32, 512),       # BN=64 path (head layer)
    (190, 96, 136),      # BN=128 path, ragged
    (1000, 2048, 4096),  # FFN down-projection shape
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("mode", ["kk", "kmn", "mnmn"])
def test_gemm_plain(ops, M, N, K, mode):
    a = _rand((M, K), 1.0, 1)
    b = _rand((N, K), 1.0, 2)
    ref = a.float() @ b.float().t()
    if mode == "kk":
        got = ops.gemm(a, b, out_dtype=F32)
    elif mode == "kmn":  # B stored [K, N]
        if N % 8:
            pytest.skip("MN-major B needs N % 8 == 0")
        got = ops.gemm(a, b.t().contiguous(), b_mn=True, out_dtype=F32)
    else:  # both stored [K, *]
        if N % 8 or M % 8:
            pytest.skip("MN-major operands need M, N % 8 == 0")
        got = ops.gemm(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True, out_dtype=F32)
    assert_close(f"gemm[{mode}] {M}x{N}x{K}", got, ref, rtol=1e-3, atol=1e-3 * math.sqrt(K))


def test_gemm_bf16_out(ops):
    a, b = _rand((512, 1024), 1.0, 3), _rand((768, 1024), 1.0, 4)
    got = ops.gemm(a, b)
    assert got.dtype == BF16
    assert_close("gemm bf16 out", got, a.float() @ b.float().t(), rtol=1e-2, atol=0.3)


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES + [(31, 64, 64), (33, 72, 64), (257, 328, 192), (5000, 4096, 128)])
def test_gemm_bf16_out_all_shapes(ops, M, N, K):
    """bf16 outputs leave through shared memory + TMA bulk stores when the pitch allows (ragged M / N are clipped by the
    tensor map); narrow or unaligned outputs take the register-store path.  Same numbers either way."""
    a, b = _rand((M, K), 1.0, 31), _rand((N, K), 0.1, 32)
    bias = _rand((N,), 1.0, 33, F32)
    res = _rand((M, N), 1.0, 34)
    mask = torch.relu(_rand((M, N), 1.0, 35))
    ref = ((a.float() @ b.float().t() + bias) * (mask.float() > 0) + res.float()).to(BF16)
    got = ops.gemm(a, b, bias=bias, relu_mask=mask, residual=res)
    assert got.dtype == BF16
    assert_close(f"gemm bf16 out {M}x{N}x{K}", got, ref.float(), rtol=1e-2, atol=3e-2)
    # into a column slice of a wider buffer (pitch > N): the neighbouring columns stay untouched
    wide = torch.full((M, N + 16), 7.0, device="cuda", dtype=BF16)
    ops.gemm(a, b, bias=bias, relu_mask=mask, residual=res, out=wide[:, 8:8 + N])
    assert torch.equal(wide[:, 8:8 + N], got)
    assert (wide[:, :8] == 7).all() and (wide[:, 8 + N:] == 7).all()
    # column sums of the stored bf16 values, accumulated in the epilogue (d b_1 of the FFN backward) on top of `colsum_out`
    if ops.gemm_fuses_colsum(M, N):
        cs = torch.full((N,), 0.5, device="cuda", dtype=F32)
        got2 = ops.gemm(a, b, bias=bias, relu_mask=mask, residual=res, colsum_out=cs)
        assert torch.equal(got2, got)
        want = got.float().sum(0) + 0.5
        assert_close(f"gemm fused colsum {M}x{N}x{K}", cs, want, rtol=1e-4, atol=1e-3 * math.sqrt(M))
    else:
        assert N < 64 or M < 32
        with pytest.raises(RuntimeError):
            ops.gemm(a, b, colsum_out=torch.zeros(N, device="cuda"))


def test_gemm_epilogue_operands_at_odd_offsets(ops):
    """mask / residual / bias views whose first element is not 16-byte aligned (column / element slices of wider buffers)
    take the scalar-load path of the epilogue instead of faulting on a vector load."""
    M, N, K = 200, 264, 128
    a, b = _rand((M, K), 1.0, 51), _rand((N, K), 0.1, 52)
    res_w, mask_w = _rand((M, N + 16), 1.0, 53), torch.relu(_rand((M, N + 16), 1.0, 54))
    bias_w = _rand((N + 4,), 1.0, 55, F32)
    for off in (1, 4, 8):
        res, mask, bias = res_w[:, off:off + N], mask_w[:, off:off + N], bias_w[(off % 4):(off % 4) + N]
        ref = (a.float() @ b.float().t() + bias) * (mask.float() > 0) + res.float()
        got = ops.gemm(a, b, bias=bias, relu_mask=mask, residual=res, out_dtype=F32)
        assert_close(f"gemm epilogue operands at element offset {off}", got, ref, rtol=1e-3, atol=1e-2)
        got16 = ops.gemm(a, b, bias=bias, relu_mask=mask, residual=res)
        assert_close(f"gemm epilogue operands at element offset {off} (bf16 out)", got16, ref, rtol=1e-2, atol=3e-2)


def test_gemm_epilogue_bias_relu_residual(ops):
    M, N, K = 777, 1032, 520
    a, b = _rand((M, K), 1.0, 5), _rand((N, K), 0.05, 6)
    bias = _rand((N,), 1.0, 7, F32)
    res = _rand((M, N), 1.0, 8)
    ref = torch.relu(a.float() @ b.float().t() + bias) + res.float()
    got = ops.gemm(a, b, bias=bias, relu=True, residual=res, out_dtype=F32)
    assert_close("gemm bias+relu+residual (fp32 out)", got, ref, rtol=1e-3, atol=1e-2)
    got16 = ops.gemm(a, b, bias=bias, relu=True, residual=res)
    assert_close("gemm bias+relu+residual (bf16 out)", got16, ref, rtol=1e-2, atol=3e-2)


def test_gemm_epilogue_relu_mask(ops):
    M, N, K = 384, 640, 256
    a, b = _rand((M, K), 1.0, 9), _rand((N, K), 0.1, 10)
    mask = torch.relu(_rand((M, N), 1.0, 11))
    ref = (a.float() @ b.float().t()) * (mask.float() > 0)
    got = ops.gemm(a, b, relu_mask=mask, out_dtype=F32)
    assert_close("gemm relu-mask", got, ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("p", [0.1, 0.6])
def test_gemm_epilogue_dropout_matches_mask_dump(ops, p):
    M, N, K = 260, 520, 128
    a, b = _rand((M, K), 1.0, 12), _rand((N, K), 0.1, 13)
    res = _rand((M, N), 1.0, 14)
    drop = (p, 1234, 77)
    keep = ops.dropout_mask(M, N, drop).float()
    frac = keep.mean().item()
    assert abs(frac - (1 - p)) < 0.01, f"keep fraction {frac} vs {1 - p}"
    ref = (a.float() @ b.float().t()) * keep / (1 - p) + res.float()
    got = ops.gemm(a, b, residual=res, dropout=drop, out_dtype=F32)
    assert_close(f"gemm dropout p={p}", got, ref, rtol=1e-3, atol=1e-2)
    # a different offset gives a different mask
    keep2 = ops.dropout_mask(M, N, (p, 1234, 78)).float()
    assert (keep2 != keep).float().mean().item() > 0.05


@pytest.mark.parametrize("split", [2, 4, 7])
def test_gemm_split_k(ops, split):
    # weight-gradient shape: small output, long reduction
    Mred, N, K = 5000, 512, 384
    dy = _rand((Mred, N), 1.0, 15)
    x = _rand((Mred, K), 1.0, 16)
    ref = dy.float().t() @ x.float()
    got = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=F32, split_k=split)
    assert_close(f"gemm split_k={split}", got, ref, rtol=1e-3, atol=0.1)
    # accumulate on top of an existing buffer
    acc = ref.clone()
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, accumulate=True)
    assert_close("gemm accumulate", acc, 2 * ref, rtol=1e-3, atol=0.2)


def test_gemm_strided_views(ops):
    # operands and outputs that are column slices of wider buffers (fused-QKV layout)
    M, K = 640, 512
    x = _rand((M, K), 1.0, 17)
    wqkv = _rand((3 * 256, K), 0.1, 18)
    full = ops.gemm(x, wqkv, out_dtype=BF16)
    ref = x.float() @ wqkv.float().t()
    assert_close("gemm fused qkv", full, ref, rtol=1e-2, atol=5e-2)
    # dgrad through a slice of the wide gradient buffer: A = dqkv[:, 256:512] (ld = 768)
    dq = _rand((M, 768), 1.0, 19)
    got = ops.gemm(dq[:, 256:512], wqkv[256:512], b_mn=True, out_dtype=F32)
    ref = dq[:, 256:512].float() @ wqkv[256:512].float()
    assert_close("gemm sliced A", got, ref, rtol=1e-3, atol=1e-2)


def test_gemm_rejects_cpu_and_bad_args(ops):
    a = torch.zeros(8, 8, dtype=BF16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a)
    a = torch.zeros(16, 12, dtype=BF16, device="cuda")  # K=12 -> pitch not a multiple of 16 bytes
    with pytest.raises(RuntimeError):
        ops.gemm(a, a)


# ------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,D", [(1, 2048), (49, 2048), (333, 1024), (62, 4096), (17, 3072), (5, 8), (11, 8192),
                                    (7, 6144), (2051, 2048)])
@pytest.mark.parametrize("xdt,ydt", [(BF16, BF16), (F32, F32), (BF16, F32)])
def test_layernorm_fwd(ops, rows, D, xdt, ydt):
    x = (_rand((rows, D), 2.0, 20, F32) + 0.5).to(xdt)
    g = _rand((D,), 1.0, 21, F32) + 1.0
    b = _rand((D,), 1.0, 22, F32)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, eps=1e-6)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, out_dtype=ydt)
    assert y.dtype == ydt
    tol = 2e-2 if ydt == BF16 else 1e-4
    assert_close(f"ln fwd {rows}x{D} {xdt}->{ydt}", y, ref, rtol=tol, atol=tol)
    assert_close("ln mean", mean, x.float().mean(-1), rtol=1e-4, atol=1e-4)
    assert_close("ln rstd", rstd, 1.0 / torch.sqrt(x.float().var(-1, unbiased=False) + 1e-6), rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("rows,D", [(1, 2048), (490, 2048), (333, 1024), (62, 4096), (3000, 2048), (11, 8192), (7, 6144),
                                    (2051, 2048), (6, 64)])
@pytest.mark.parametrize("dt", [BF16, F32])
def test_layernorm_bwd(ops, rows, D, dt):
    x = (_rand((rows, D), 2.0, 23, F32) + 0.5).to(dt)
    g = _rand((D,), 1.0, 24, F32) + 1.0
    b = _rand((D,), 1.0, 25, F32)
    dy = _rand((rows, D), 1.0, 26, F32).to(dt)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gr, br, eps=1e-6).backward(dy.float())
    _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, out_dtype=dt)
    dx, dx_drop, dgamma, dbeta = ops.layernorm_bwd(dy, x, g, mean, rstd)
    assert dx_drop is None
    tol = 2e-2 if dt == BF16 else 1e-4
    assert_close(f"ln bwd dx {rows}x{D} {dt}", dx, xr.grad, rtol=tol, atol=tol)
    assert_close("ln bwd dgamma", dgamma, gr.grad, rtol=1e-3, atol=1e-3 * math.sqrt(rows) + 1e-3)
    assert_close("ln bwd dbeta", dbeta, br.grad, rtol=1e-3, atol=1e-3 * math.sqrt(rows) + 1e-3)


def test_layernorm_bwd_dropout_output(ops):
    rows, D, p = 200, 2048, 0.2
    x, dy = _rand((rows, D), 1.0, 27), _rand((rows, D), 1.0, 28)
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    _, mean, rstd = ops.layernorm_fwd(x, g, b)
    drop = (p, 99, 5)
    dx, dx_drop, _, _ = ops.layernorm_bwd(dy, x, g, mean, rstd, dropout=drop)
    keep = ops.dropout_mask(rows, D, drop).float()
    assert_close("ln bwd dx_drop", dx_drop, dx.float() * keep / (1 - p), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("rows", [3, 490, 2051])
@pytest.mark.parametrize("dy_dt", [BF16, F32])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_layernorm_bwd_dxsum_and_mixed_dtypes(ops, rows, dy_dt, p):
    """dy fp32 | bf16 with bf16 x (the last block's LayerNorm hands an fp32 gradient back), the column sums of the
    tensor the preceding Linear receives (its bias gradient) and the dropout copy, ragged row groups included."""
    D = 2048
    x = _rand((rows, D), 1.5, 40)
    g = _rand((D,), 1.0, 41, F32) + 1.0
    b = _rand((D,), 1.0, 42, F32)
    dy = _rand((rows, D), 1.0, 43, F32).to(dy_dt)
    xr = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), g, b, eps=1e-6).backward(dy.float())
    _, mean, rstd = ops.layernorm_fwd(x, g, b)
    drop = (p, 7, 3) if p > 0 else (0.0, 0, 0)
    dx, dx_drop, dgamma, dbeta, dxsum = ops.layernorm_bwd(dy, x, g, mean, rstd, dropout=drop, want_dxsum=True)
    assert dx.dtype == BF16
    assert_close("ln bwd mixed dx", dx, xr.grad, rtol=2e-2, atol=2e-2)
    src = dx
    if p > 0:
        keep = ops.dropout_mask(rows, D, drop).float()
        assert_close("ln bwd mixed dx_drop", dx_drop, dx.float() * keep / (1 - p), rtol=1e-2, atol=1e-3)
        src = dx_drop
    else:
        assert dx_drop is None
    assert_close("ln bwd dxsum", dxsum, src.float().sum(0), rtol=1e-3, atol=1e-3 * math.sqrt(rows) + 1e-3)


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, W, L, H, dk, bias, scale, keep=None, p=0.0):
    q, k, v = qkv.float().view(W, L, 3, H, dk).permute(2, 0, 3, 1, 4)  # [W,H,L,dk]
    s = (q * scale) @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias.unsqueeze(0)
    prob = torch.softmax(s, dim=-1)
    if keep is not None:
        prob = prob * keep.view(W, H, L, L) / (1 - p)
    o = prob @ v
    return o.transpose(1, 2).reshape(W * L, H * dk), prob


ATTN_CASES = [(3, 49, 8, 256), (5, 19, 8, 256), (2, 81, 8, 256), (4, 17, 8, 256), (1, 1, 2, 256), (3, 33, 4, 128),
              (2, 96, 2, 64), (300, 49, 8, 256), (2, 80, 3, 256), (2, 16, 2, 256)]


@pytest.mark.parametrize("W,L,H,dk", ATTN_CASES)
@pytest.mark.parametrize("use_bias", [False, True])
def test_attn_fwd(ops, W, L, H, dk, use_bias):
    qkv = _rand((W * L, 3 * H * dk), 1.0, 30)
    bias = None
    if use_bias:
        bias = _rand((H, L, L), 0.5, 31, F32)
        bias[:, 0, :] = 0
        bias[:, :, 0] = 0
    scale = 1.0 / math.sqrt(dk)
    ref_o, ref_p = _attn_ref(qkv, W, L, H, dk, bias, scale)
    out, probs = ops.attn_fwd(qkv, W, L, H, dk, bias, scale, return_probs=True)
    assert_close(f"attn fwd probs W{W} L{L} H{H} dk{dk}", probs, ref_p, rtol=1e-2, atol=2e-3)
    assert_close(f"attn fwd out W{W} L{L} H{H} dk{dk}", out, ref_o, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("W,L,H,dk", ATTN_CASES)
def test_attn_bwd(ops, W, L, H, dk):
    qkv = _rand((W * L, 3 * H * dk), 1.0, 32)
    bias = _rand((H, L, L), 0.5, 33, F32)
    bias[:, 0, :] = 0
    bias[:, :, 0] = 0
    dout = _rand((W * L, H * dk), 1.0, 34)
    scale = 1.0 / math.sqrt(dk)
    qr = qkv.float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    o, _ = _attn_ref(qr, W, L, H, dk, br, scale)
    o.backward(dout.float())
    dqkv, dbias = ops.attn_bwd(qkv, dout, W, L, H, dk, bias, scale, need_dbias=True)
    gmax = qr.grad.abs().max().item()
    assert_close(f"attn bwd dqkv W{W} L{L} H{H} dk{dk}", dqkv, qr.grad, rtol=3e-2, atol=3e-2 * max(gmax, 1e-3))
    ref_db = br.grad.clone()
    ref_db[:, 0, :] = 0
    ref_db[:, :, 0] = 0
    assert_close("attn bwd dbias", dbias, ref_db, rtol=3e-2, atol=3e-2 * max(ref_db.abs().max().item(), 1e-3))


@pytest.mark.parametrize("W,L,H,dk", [(4, 49, 8, 256), (3, 19, 8, 256), (2, 81, 8, 256)])
def test_attn_dropout_fwd_bwd(ops, W, L, H, dk):
    p = 0.2
    drop = (p, 4242, 3)
    qkv = _rand((W * L, 3 * H * dk), 1.0, 35)
    dout = _rand((W * L, H * dk), 1.0, 36)
    scale = 1.0 / math.sqrt(dk)
    keep = ops.dropout_mask(W * H * L, L, drop).float()
    qr = qkv.float().requires_grad_(True)
    ref_o, ref_p = _attn_ref(qr, W, L, H, dk, None, scale, keep, p)
    ref_o.backward(dout.float())
    out, probs = ops.attn_fwd(qkv, W, L, H, dk, None, scale, dropout=drop, return_probs=True)
    assert_close("attn dropout probs", probs, ref_p, rtol=1e-2, atol=2e-3)
    assert_close("attn dropout out", out, ref_o, rtol=2e-2, atol=2e-2)
    dqkv, _ = ops.attn_bwd(qkv, dout, W, L, H, dk, None, scale, dropout=drop)
    assert_close("attn dropout dqkv", dqkv, qr.grad, rtol=3e-2, atol=3e-2 * qr.grad.abs().max().item())


def test_relbias_gather_scatter(ops):
    T, H, n, L = 245, 8, 48, 49
    table = _rand((T, H), 0.02, 37, F32)
    g = torch.Generator(device="cuda").manual_seed(38)
    index = torch.randint(0, T, (n, n), generator=g, device="cuda")
    dense = ops.relbias_gather(table, index, L)
    ref = torch.zeros(H, L, L, device="cuda")
    ref[:, 1:, 1:] = table[index.reshape(-1)].reshape(n, n, H).permute(2, 0, 1)
    assert torch.equal(dense, ref)
    # sliced index (UCF: 18 of 32)
    dense18 = ops.relbias_gather(table, index, 19)
    ref18 = torch.zeros(H, 19, 19, device="cuda")
    ref18[:, 1:, 1:] = table[index[:18, :18].reshape(-1)].reshape(18, 18, H).permute(2, 0, 1)
    assert torch.equal(dense18, ref18)
    dd = _rand((H, L, L), 1.0, 39, F32)
    dt = ops.relbias_scatter(dd, index, T)
    tr = table.clone().requires_grad_(True)
    (tr[index.reshape(-1)].reshape(n, n, H).permute(2, 0, 1) * dd[:, 1:, 1:]).sum().backward()
    assert_close("relbias scatter", dt, tr.grad, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------ glue
@pytest.mark.parametrize("W,L0,D", [(5, 48, 2048), (3, 16, 2048), (2, 80, 1024), (1, 18, 2048), (7, 1, 64)])
@pytest.mark.parametrize("dt", [F32, BF16])
def test_cls_prepend(ops, W, L0, D, dt):
    x = _rand((W, L0, D), 1.0, 40, F32).abs().to(dt)
    ref = torch.cat([x.float().mean(1, keepdim=True), x.float()], 1)
    out = ops.cls_prepend_fwd(x)
    assert_close(f"cls prepend {W}x{L0}x{D} {dt}", out, ref, rtol=1e-2, atol=1e-2)
    # learned token + position table
    cls = _rand((D,), 1.0, 41, F32)
    pos = _rand((100, D), 1.0, 42, F32)
    if L0 + 1 <= 100:
        ref2 = torch.cat([cls.view(1, 1, D).expand(W, 1, D), x.float()], 1) + pos[: L0 + 1]
        out2 = ops.cls_prepend_fwd(x, cls=cls, pos=pos)
        assert_close("cls prepend learned+pos", out2, ref2, rtol=1e-2, atol=2e-2)
    # backward
    g = _rand((W, L0 + 1, D), 1.0, 43)
    dx, dcls, dpos = ops.cls_prepend_bwd(g, cls_learned=False, need_dx=True, need_dpos=False)
    refdx = g.float()[:, 1:] + g.float()[:, :1] / L0
    assert_close("cls prepend bwd dx", dx, refdx, rtol=1e-5, atol=1e-5)
    dx2, dcls2, dpos2 = ops.cls_prepend_bwd(g, cls_learned=True, need_dx=True, need_dpos=True)
    assert_close("cls prepend bwd dx (learned)", dx2, g.float()[:, 1:], rtol=1e-5, atol=1e-5)
    assert_close("cls prepend bwd dcls", dcls2, g.float()[:, 0].sum(0), rtol=1e-4, atol=1e-3)
    assert_close("cls prepend bwd dpos", dpos2, g.float().sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("R,C", [(64, 64), (2048, 4096), (100, 37), (3027, 520), (1, 9)])
def test_cast_transposed_is_bit_exact(ops, R, C):
    x = _rand((R, C), 1.0, 41, F32)
    assert torch.equal(ops.cast_to_bf16_transposed(x), x.to(BF16).t().contiguous())
    # source = row slice of a taller matrix, destination = column slice of a wider one (the fused [D, 3*H*dk] QKV^T buffer)
    wide = torch.full((C, 3 * R + 8), 3.0, device="cuda", dtype=BF16)
    ops.cast_to_bf16_transposed(x, out=wide[:, R:2 * R])
    assert torch.equal(wide[:, R:2 * R], x.to(BF16).t())
    assert (wide[:, :R] == 3).all() and (wide[:, 2 * R:] == 3).all()


def test_casts_colsum_scale(ops):
    x = _rand((1000, 333), 3.0, 44, F32)
    assert torch.equal(ops.cast_to_bf16(x), x.to(BF16))
    xb = x.to(BF16)
    assert torch.equal(ops.cast_to_f32(xb), xb.float())
    big = _rand((62720 // 8, 4096), 1.0, 45)
    assert_close("colsum", ops.colsum(big), big.float().sum(0), rtol=1e-4, atol=5e-2)
    odd = _rand((77, 31), 1.0, 46)
    assert_close("colsum odd", ops.colsum(odd), odd.float().sum(0), rtol=1e-4, atol=1e-3)
    sl = big[:, 512:1024]
    assert_close("colsum slice", ops.colsum(sl), sl.float().sum(0), rtol=1e-4, atol=5e-2)
    s = torch.tensor([0.37], device="cuda")
    assert_close("scale", ops.scale_by_device_scalar(x, s), x * 0.37, rtol=1e-6, atol=1e-6)
    d = ops.dropout_apply(xb[:, :328].contiguous(), (0.3, 5, 6))
    keep = ops.dropout_mask(1000, 328, (0.3, 5, 6)).float()
    assert_close("dropout apply", d, xb[:, :328].float() * keep / 0.7, rtol=1e-2, atol=1e-3)


def test_adagrad_matches_torch(ops):
    n = 100003
    p = _rand((n,), 1.0, 47, F32)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adagrad([pr], lr=1e-2, weight_decay=1e-3)
    st = torch.zeros_like(p)
    for i in range(3):
        g = _rand((n,), 1.0, 48 + i, F32)
        pr.grad = g.clone()
        opt.step()
        ops.adagrad_step(p, g, st, lr=1e-2, weight_decay=1e-3, eps=1e-10)
    assert_close("adagrad", p, pr.detach(), rtol=1e-5, atol=1e-6)
    # gradient-norm clipping coefficient, no host sync
    gs = [_rand((1000,), 3.0, 60, F32), _rand((77, 5), 3.0, 61, F32)]
    coef = ops.grad_clip_coef(gs, 10.0)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs)).item()
    assert abs(coef.item() - min(1.0, 10.0 / (total + 1e-6))) < 1e-6
    assert ops.grad_clip_coef(gs, 1e9).item() == 1.0


# ------------------------------------------------------------------------------------------ heads
@pytest.mark.parametrize("C,sigmoid", [(2, False), (1, True)])
@pytest.mark.parametrize("n", [1, 37, 1280])
def test_head_tail(ops, C, sigmoid, n):
    K1 = 512
    h1 = torch.relu(_rand((n, K1), 1.0, 50))
    W2, b2 = _rand((32, K1), 0.05, 51, F32), _rand((32,), 0.1, 52, F32)
    W3, b3 = _rand((C, 32), 0.3, 53, F32), _rand((C,), 0.1, 54, F32)
    W2r, b2r, W3r, b3r = (t.clone().requires_grad_(True) for t in (W2, b2, W3, b3))
    h1r = h1.float().requires_grad_(True)
    h2r = h1r @ W2r.t() + b2r
    z = h2r @ W3r.t() + b3r
    ref = torch.sigmoid(z) if sigmoid else torch.softmax(z, -1)
    dout = _rand((n, C), 1.0, 55, F32)
    ref.backward(dout)
    h2, out = ops.head_tail_fwd(h1, W2, b2, W3, b3, sigmoid)
    assert_close(f"head tail h2 C{C} n{n}", h2, h2r, rtol=1e-2, atol=1e-2)
    assert_close(f"head tail out C{C} n{n}", out, ref, rtol=1e-2, atol=3e-3)
    dh2, dW3, db3 = ops.head_tail_bwd(dout, out, h2, W3, sigmoid)
    # reference gradients taken at the bf16-rounded h2 the kernel saved
    zz = (h2.float() @ W3.t() + b3).requires_grad_(True)
    rr = torch.sigmoid(zz) if sigmoid else torch.softmax(zz, -1)
    rr.backward(dout)
    assert_close("head tail dh2", dh2, zz.grad @ W3, rtol=2e-2, atol=2e-3)
    assert_close("head tail dW3", dW3, zz.grad.t() @ h2.float(), rtol=1e-2, atol=1e-2)
    assert_close("head tail db3", db3, zz.grad.sum(0), rtol=1e-2, atol=1e-3)


# ------------------------------------------------------------------------------------------ losses
def _mil_ref(y, B, P, T, lambda1, spar_start):
    part = y.view(2 * B, P, T).mean(-1)
    top = part.max(-1)[0]
    nor, abn = top[:B], top[B:]
    err = sum(torch.relu(1 - abn + nor[i]).sum() for i in range(B)) / B ** 2
    spar = y.reshape(-1)[spar_start:].mean()
    return err + lambda1 * spar, err, spar


@pytest.mark.parametrize("B,P,T", [(40, 16, 1), (40, 32, 1), (8, 16, 7), (1, 1, 1), (3, 5, 2)])
def test_mil_loss(ops, B, P, T):
    g = torch.Generator(device="cuda").manual_seed(60)
    y = torch.rand(2 * B * P * T, generator=g, device="cuda")
    spar_start = B if T == 1 else B * P * T
    yr = y.clone().requires_grad_(True)
    loss, err, spar = _mil_ref(yr, B, P, T, 0.01, spar_start)
    loss.backward()
    out3, top_idx, dy = ops.mil_loss(y, B, P, T, 1, 0.01, spar_start)
    assert_close("mil loss", out3, torch.stack([loss, err, spar]).detach(), rtol=1e-5, atol=1e-6)
    assert_close("mil grad", dy, yr.grad, rtol=1e-5, atol=1e-7)
    ref_idx = y.view(2 * B, P, T).mean(-1).argmax(-1).int()
    part = y.view(2 * B, P, T).mean(-1)
    if T == 1:  # unique maxima with prob. 1 -> indices are bit-exact
        assert torch.equal(top_idx[:, 0], ref_idx)
    else:
        assert torch.equal(part.gather(1, top_idx.long()), part.max(-1, keepdim=True)[0])


def test_mil_loss_ties_pick_lowest_index_and_topk(ops):
    B, P = 2, 8
    y = torch.zeros(2 * B * P, device="cuda")
    y[3] = y[5] = 0.9  # tie inside bag 0
    out3, idx, dy = ops.mil_loss(y, B, P, 1, 1, 0.0, B)
    assert idx[0, 0].item() == 3
    y = torch.rand(2 * B * P, device="cuda")
    out3, idx, dy = ops.mil_loss(y, B, P, 1, 3, 0.01, B)
    top3 = y.view(2 * B, P).topk(3, dim=-1)
    assert torch.equal(idx.long().sort(-1)[0], top3.indices.sort(-1)[0])
    yr = y.clone().requires_grad_(True)
    bag = yr.view(2 * B, P).topk(3, dim=-1).values.mean(-1)
    err = sum(torch.relu(1 - bag[B:] + bag[i]).sum() for i in range(B)) / B ** 2
    (err + 0.01 * yr[B:].mean()).backward()
    assert_close("mil topk grad", dy, yr.grad, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n", [1, 1280, 2560])
def test_soft_ce_and_bce(ops, n):
    g = torch.Generator(device="cuda").manual_seed(61)
    probs = torch.softmax(torch.randn(n, 2, generator=g, device="cuda"), -1)
    m = torch.rand(n, generator=g, device="cuda")
    lab = torch.stack([1 - m, m], -1)
    pr = probs.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(pr, lab)
    ref.backward()
    out1, dp = ops.soft_ce_loss(probs, lab)
    assert_close("soft ce", out1, ref.detach().view(1), rtol=1e-5, atol=1e-6)
    assert_close("soft ce grad", dp, pr.grad, rtol=1e-4, atol=1e-7)
    T = 7
    s = torch.rand(n * T, generator=g, device="cuda") * 0.98 + 0.01
    sr = s.clone().requires_grad_(True)
    o = sr.view(n, T).mean(-1)
    refb = torch.mean(-0.2 * lab[:, 0] * torch.log(1 - o + 1e-8) - 2.0 * lab[:, 1] * torch.log(o + 1e-8))
    refb.backward()
    ob, ds = ops.bce_loss(s, lab, T, 0.2, 2.0)
    assert_close("bce", ob, refb.detach().view(1), rtol=1e-5, atol=1e-6)
    assert_close("bce grad", ds, sr.grad, rtol=1e-4, atol=1e-8)


def test_threshold_labels_bit_exact(ops):
    s = torch.rand(100000, device="cuda")
    s[:10] = 0.65
    out = ops.threshold_labels(s, 0.65)
    assert torch.equal(out, torch.where(s > 0.65, s, torch.zeros_like(s)))


# ------------------------------------------------------------------------------------------ CLS-query attention
@pytest.mark.parametrize("W,L,H,dk", [(3, 49, 8, 256), (5, 19, 8, 256), (2, 81, 8, 256), (4, 17, 2, 64), (1, 1, 1, 128),
                                      (300, 49, 8, 256), (2, 96, 2, 256)])
@pytest.mark.parametrize("p", [0.0, 0.2])
def test_attn_cls_matches_row0_of_full_attention(ops, W, L, H, dk, p):
    HD = H * dk
    qkv = _rand((W * L, 3 * HD), 1.0, 70)
    scale = 1.0 / math.sqrt(dk)
    drop = (p, 777, 9) if p > 0 else ops.NO_DROPOUT
    keep = ops.dropout_mask(W * H * L, L, drop).float() if p > 0 else None
    qr = qkv.float().requires_grad_(True)
    ref_o, _ = _attn_ref(qr, W, L, H, dk, None, scale, keep, p)
    ref_cls = ref_o.view(W, L, HD)[:, 0, :]
    dout = _rand((W, HD), 1.0, 71)
    ref_cls.backward(dout.float())
    q_cls = qkv.view(W, L, 3 * HD)[:, 0, :HD].contiguous()
    kv = qkv[:, HD:].contiguous()
    o = ops.attn_cls_fwd(q_cls, kv, W, L, H, dk, scale, drop)
    assert_close(f"attn cls fwd W{W} L{L} H{H} dk{dk} p{p}", o, ref_cls, rtol=2e-2, atol=2e-2)
    dq, dkv = ops.attn_cls_bwd(q_cls, kv, dout, W, L, H, dk, scale, drop)
    g = qr.grad.view(W, L, 3 * HD)
    gmax = max(g.abs().max().item(), 1e-3)
    assert_close("attn cls bwd dq", dq, g[:, 0, :HD], rtol=3e-2, atol=3e-2 * gmax)
    # rows 1.. of the q-gradient are zero in the reference (only the CLS query was used)
    assert g[:, 1:, :HD].abs().max().item() == 0 if L > 1 else True
    assert_close("attn cls bwd dkv", dkv.view(W, L, 2 * HD), g[:, :, HD:], rtol=3e-2, atol=3e-2 * gmax)


def test_add_rows(ops):
    a = _rand((40, 49 * 64), 1.0, 72)
    b = _rand((40, 64), 1.0, 73)
    ref = a.float().clone()
    ref[:, :64] += b.float()
    ops.add_rows_(a[:, :64], b)
    assert_close("add rows", a, ref, rtol=1e-2, atol=1e-2)
