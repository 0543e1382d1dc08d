"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded inputs.

TEST INFRASTRUCTURE.  Runs only in the build container (the reference checkout is not shipped to the GPU
box); its outputs are committed so that every later test — CPU oracle and CUDA path — can be checked
against what the reference itself computes.

    python oracle/make_golden.py            # rewrites tests/golden/

For every case the fixture stores: the constructor kwargs, the full state_dicts (fp32, exactly as
`state_dict()` returns them, i.e. also proving key names / shapes / buffer dtypes), the inputs, the
forward outputs (eval mode), the loss values of the reference's own loss functions and the gradients of
every parameter and of the input.  Cases are kept small (d_model 64..128) so the fixtures stay ~MBs.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    if not REF.exists():
        raise SystemExit("reference checkout not present; golden vectors can only be regenerated in the build container")
    for name in ("h5py", "matplotlib", "matplotlib.pyplot"):  # imported by Train/*.py, unused by the hot path
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    # drop any non-reference `models` / `utils` package from the import cache so the reference's packages resolve
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    sys.path.insert(0, str(REF / "Train"))
    sys.path.insert(0, str(REF))
    from models.Encoder import Encoder
    from models.Classifier import Classifier
    from models.Regressor import Regressor
    ltn = importlib.import_module("temporal_transformer_shanghaitech")
    stn = importlib.import_module("spatio_transformer_shanghaitech")
    milce = importlib.import_module("spatio_transformer_MIL_CE")
    return Encoder, Classifier, Regressor, ltn, stn, milce


def feats(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).abs()  # I3D features are post-ReLU: non-negative


def grads_of(module):
    return {k: (p.grad.clone() if p.grad is not None else None) for k, p in module.named_parameters()}


def run_ltn_case(name, enc_kw, B, P, T, N, seed, Encoder, Classifier, ltn, extra_L0=()):
    torch.manual_seed(seed)
    enc = Encoder(**enc_kw)
    cls = Classifier(enc_kw["d_model"], 0.6, weight_init=True)
    D = enc_kw["d_model"]
    enc.eval(); cls.eval()  # eval: dropout off, deterministic
    x = feats((2 * B * P, T * N, D), seed + 1).requires_grad_(True)
    g = torch.Generator().manual_seed(seed + 2)
    clip_pseudo = (torch.rand(B, P * T, 1, generator=g) > 0.5).float() * torch.rand(B, P * T, 1, generator=g)

    args = types.SimpleNamespace(batch_size=B, part_num=P, part_len=T, lambda_1=0.01, lambda_MIL=1.0, lambda_CE=0.8)
    # label prep exactly as Train/temporal_transformer_shanghaitech.py:103-112 (on CPU)
    norm_labs = torch.zeros([B, P, 2]); norm_labs[:, :, 0] += 1
    ab = clip_pseudo.view([B, P, T]).mean(-1).view([B, P, 1])
    ab2 = torch.zeros([B, P, 2]); ab2[:, :, 1] = ab[:, :, 0]; ab2[:, :, 0] = 1 - ab2[:, :, 1]
    clip_labs = torch.cat([norm_labs, ab2], 0).view([2 * B * P, -1])

    enc_out, attn_list, v_list = enc(x, return_attn_v=True)
    f = enc_out[:, 0, :].float().view([2 * B, P, D])
    outputs = cls(f).view([2 * B * P, -1])
    score = outputs[:, 1]
    ce = ltn.get_CE_loss(args, outputs, clip_labs)
    mil, err, l1 = ltn.get_MIL_loss(args, score)
    loss = args.lambda_MIL * mil + args.lambda_CE * ce
    loss.backward()
    topk_idx = score.view(2 * B, P).max(-1)[1]
    case = dict(kind="ltn", enc_kwargs=enc_kw, B=B, P=P, T=T, N=N,
                enc_state=enc.state_dict(), cls_state=cls.state_dict(),
                x=x.detach().clone(), clip_pseudo=clip_pseudo, clip_labs=clip_labs,
                enc_out=enc_out.detach(), attn0=attn_list[0].detach(), v0=v_list[0].detach(),
                probs=outputs.detach(), ce=ce.detach(), mil=mil.detach(), err=err.detach(), spar=l1.detach(),
                loss=loss.detach(), topk_idx=topk_idx, x_grad=x.grad.clone(),
                enc_grads=grads_of(enc), cls_grads=grads_of(cls),
                thr_labels=torch.where(score.detach() > 0.5, score.detach(), torch.zeros_like(score.detach())))
    # variable-length windows through the same weights (short trailing windows of the labelling loop)
    with torch.no_grad():
        case["var_L"] = {}
        for L0 in extra_L0:
            xv = feats((3, L0, D), seed + 10 + L0)
            case["var_L"][L0] = dict(x=xv, enc_out=enc(xv), probs=cls(enc(xv)[:, 0, :]))
    torch.save(case, OUT / f"{name}.pt")
    print(f"{name}: loss={loss.item():.6f} enc_out={tuple(enc_out.shape)}")


def run_stn_case(name, enc_kw, hidden_kw, B, P, T, N, seed, Encoder, Regressor, stn, milce):
    torch.manual_seed(seed)
    enc = Encoder(**enc_kw)
    reg = Regressor(enc_kw["d_model"], 0.6, weight_init=True)
    D = enc_kw["d_model"]
    enc.eval(); reg.eval()
    x = feats((2 * B * P * T, N, D), seed + 1).requires_grad_(True)
    args = types.SimpleNamespace(batch_size=B, part_num=P, part_len=T, lambda_1=0.01, lambda_normal=0.2,
                                 lambda_abnormal=2.0)
    enc_out = enc(x)
    f = enc_out[:, 0, :].float().view([2 * B, P * T, D])
    outputs = reg(f).view([2 * B, P * T, 1])
    mil, err, l1 = stn.get_MIL_loss(args, outputs)
    # BCE on per-part mean scores against soft labels (Train/spatio_transformer_MIL_CE.py:176-181)
    g = torch.Generator().manual_seed(seed + 3)
    m = torch.rand(2 * B, P, generator=g)
    labs = torch.stack([1 - m, m], -1)
    part_scores = torch.mean(outputs.view([2 * B, P, T]), dim=-1)
    bce = milce.get_BCE_loss(args, part_scores, labs)
    loss = mil + 0.5 * bce
    loss.backward()
    case = dict(kind="stn", enc_kwargs=enc_kw, B=B, P=P, T=T, N=N,
                enc_state=enc.state_dict(), reg_state=reg.state_dict(), x=x.detach().clone(),
                enc_out=enc_out.detach(), scores=outputs.detach(), mil=mil.detach(), err=err.detach(),
                spar=l1.detach(), bce=bce.detach(), bce_labs=labs, loss=loss.detach(), x_grad=x.grad.clone(),
                enc_grads=grads_of(enc), reg_grads=grads_of(reg))
    torch.save(case, OUT / f"{name}.pt")
    print(f"{name}: loss={loss.item():.6f} enc_out={tuple(enc_out.shape)}")


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    Encoder, Classifier, Regressor, ltn, stn, milce = import_reference()
    torch.set_num_threads(4)
    # 1. LTN SHT-like: 3-D rel-pos bias (ws 4, wd 3 => 48 tokens), MHA-LN + FFN-LN, H*dk != d_model, default init
    run_ltn_case("ltn_relpe", dict(n_layers=2, n_head=2, d_k=64, d_v=64, d_model=64, d_inner=128,
                                   MHA_layerNorm=True, FFN_layerNorm=True, weight_init=False, relative_pe=True,
                                   window_size=4, window_depth=3),
                 B=2, P=4, T=3, N=16, seed=0, Encoder=Encoder, Classifier=Classifier, ltn=ltn, extra_L0=(16, 32))
    # 2. LTN UCF-like: 9 patches x part_len 2 with the window_size-4 index sliced to 18x18
    run_ltn_case("ltn_ucf_sliced", dict(n_layers=1, n_head=2, d_k=64, d_v=64, d_model=128, d_inner=256,
                                        MHA_layerNorm=True, FFN_layerNorm=True, weight_init=False, relative_pe=True,
                                        window_size=4, window_depth=2),
                 B=3, P=4, T=2, N=9, seed=10, Encoder=Encoder, Classifier=Classifier, ltn=ltn, extra_L0=(9,))
    # 3. STN SHT-like: no rel-pos, no MHA-LN, xavier init, odd FFN width (3027-style, not a multiple of 8)
    run_stn_case("stn_plain", dict(n_layers=2, n_head=2, d_k=64, d_v=64, d_model=128, d_inner=203,
                                   MHA_layerNorm=False, FFN_layerNorm=True, weight_init=True),
                 None, B=2, P=3, T=2, N=16, seed=20, Encoder=Encoder, Regressor=Regressor, stn=stn, milce=milce)
    # 4. STN with the 2-D rel-pos bias, learned CLS token, absolute position encoding and input LayerNorm
    run_stn_case("stn_relpe2d_cls_pos", dict(n_layers=1, n_head=1, d_k=64, d_v=64, d_model=64, d_inner=128,
                                             MHA_layerNorm=True, FFN_layerNorm=False, weight_init=True,
                                             CLS_learned=True, position_encoding=True, max_position_tokens=40,
                                             relative_pe_2D=True, window_size=4, input_layerNorm=True),
                 None, B=2, P=2, T=2, N=16, seed=30, Encoder=Encoder, Regressor=Regressor, stn=stn, milce=milce)
    # 5. integer index buffers at the shipped configurations (bit-exact check of the oracle's restatement)
    from models.MultiHeadAttention import MultiHeadAttention
    idx = {}
    for ws, wd in ((4, 3), (4, 2), (4, 5), (3, 3)):
        idx[f"3d_ws{ws}_wd{wd}"] = MultiHeadAttention(1, 8, 8, 8, relative_pe=True, window_size=ws,
                                                      window_depth=wd).relative_position_index.clone()
    for ws in (4, 3):
        idx[f"2d_ws{ws}"] = MultiHeadAttention(1, 8, 8, 8, relative_pe_2D=True,
                                               window_size=ws).relative_position_index.clone()
    torch.save(idx, OUT / "relpos_index.pt")
    print("relpos_index:", {k: tuple(v.shape) for k, v in idx.items()})


if __name__ == "__main__" and "--sampler" not in sys.argv:
    main()


def golden_sampler():
    """chosen-index golden vectors from the reference's own `sample_feat` (utils/load_dataset.py:56-88)."""
    import numpy as np
    import_reference()
    ld = importlib.import_module("utils.load_dataset")
    cases = []
    for sample in ("uniform", "random"):
        for (n, P, T) in ((200, 16, 3), (17, 16, 3), (50, 32, 2), (19, 16, 3), (400, 16, 5), (33, 4, 7)):
            for seed in (0, 1):
                self = types.SimpleNamespace(sample=sample, part_len=T, part_num=P)
                feat = np.arange(n, dtype=np.float32).reshape(n, 1)   # feature value == clip index
                np.random.seed(seed)
                chosen_feat, labs = ld.SH_Train_Origin_Dataset.sample_feat(self, feat, None, vid_type="Abnormal")
                cases.append(dict(sample=sample, n=n, P=P, T=T, seed=seed, chosen=chosen_feat[:, 0].astype(np.int64)))
    torch.save(cases, OUT / "window_sampler.pt")
    print("window_sampler:", len(cases), "cases")


if __name__ == "__main__" and "--sampler" in sys.argv:
    golden_sampler()
