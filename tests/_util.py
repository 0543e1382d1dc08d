"""Comparison helper that prints enough to debug a layout bug from a log file."""
import torch


def report(name, got, ref, rtol, atol):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    diff = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = diff > tol
    nbad = int(bad.sum())
    finite = bool(torch.isfinite(got).all())
    msg = (f"{name}: shape={tuple(got.shape)} max_abs={diff.max().item():.4e} "
           f"ref_absmax={ref.abs().max().item():.4e} mean_abs={diff.mean().item():.4e} "
           f"bad={nbad}/{got.numel()} finite={finite}")
    if nbad or not finite:
        idx = bad.nonzero()[:8].tolist()
        vals = [(i, float(got[tuple(i)]), float(ref[tuple(i)])) for i in idx]
        msg += f" first_bad(idx,got,ref)={vals}"
        if got.dim() == 2:
            rows_bad = bad.any(dim=1).nonzero().flatten()
            cols_bad = bad.any(dim=0).nonzero().flatten()
            msg += (f" bad_rows[{len(rows_bad)}]={rows_bad[:16].tolist()} "
                    f"bad_cols[{len(cols_bad)}]={cols_bad[:16].tolist()}")
    print(msg, flush=True)
    return nbad == 0 and finite, msg


def assert_close(name, got, ref, rtol=2e-2, atol=2e-2):
    ok, msg = report(name, got, ref, rtol, atol)
    assert ok, msg
