"""Script-level drop-in check (SURVEY.md §4 tier): the batch-loop body of Train/temporal_transformer_shanghaitech.py:99-142
- the script's OWN get_CE_loss / get_MIL_loss (:21-36), its nn.DataParallel wrap (:76-78), clip_grad_norm_ (:139-141)
and torch.optim.Adagrad with two parameter groups (:83-85) - runs unmodified against `dropin/models` on CUDA: the
modules are imported through the reference's import lines (`from models.Encoder import Encoder`), their outputs are
ordinary autograd tensors consumed by plain torch ops, and the stock optimizer updates their parameters.
The reference tree does not exist on the GPU box, so the loop body is restated here line by line (it is 40 lines of
script code, not library code); the first iteration's loss is checked against the CPU oracle on the same weights and data."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOOP = r'''
import json, sys, types
import torch, torch.nn as nn, torch.nn.functional as F
from models.Encoder import Encoder            # Train/temporal_transformer_shanghaitech.py:16-18
from models.Classifier import Classifier
import models.Encoder as _m
assert "dropin" in _m.__file__, _m.__file__

args = types.SimpleNamespace(batch_size=4, part_num=4, part_len=3, n_patch=16, d_model=256, n_hidden=512, n_layers=2,
    n_head=2, d_k=64, d_v=64, MHA_attn_dropout=0.0, MHA_fc_dropout=0.0, MHA_layerNorm=True, FFN_dropout=0.0,
    FFN_layerNorm=True, position_dropout=0.1, encoder_weight_init=False, position_encoding=False, CLS_learned=False,
    max_position_tokens=100, relative_position_encoding=True, window_size=4, conv_patch=False, classifier_dropout=0.0,
    classifier_weight_init=True, data_parallel=True, temporal_only=False, lambda_1=0.01, lambda_MIL=1.0, lambda_CE=0.8,
    clip_grad=True, lr_encoder=1e-4, lr_classifier=1e-2, weight_decay=1e-3)

def get_CE_loss(args, outputs, labs):          # :21-23
    loss = F.cross_entropy(outputs, labs)
    return loss

def get_MIL_loss(args, y_pred):                # :25-36
    topk_pred = torch.max(y_pred.view([args.batch_size * 2, args.part_num]), dim=-1, keepdim=False)[0]
    nor_max = topk_pred[:args.batch_size]
    abn_max = topk_pred[args.batch_size:]
    err = 0
    for i in range(args.batch_size):
        err += torch.sum(F.relu(1 - abn_max + nor_max[i]))
    err = err / (args.batch_size) ** 2
    abn_pred = y_pred[args.batch_size:]
    spar_l1 = torch.mean(abn_pred)
    loss = err + args.lambda_1 * spar_l1
    return loss, err, spar_l1

torch.manual_seed(0)
temporal_model = Encoder(n_layers=args.n_layers, n_head=args.n_head, d_k=args.d_k, d_v=args.d_v, d_model=args.d_model,
                         d_inner=args.n_hidden, MHA_attn_dropout=args.MHA_attn_dropout, MHA_fc_dropout=args.MHA_fc_dropout,
                         MHA_layerNorm=args.MHA_layerNorm, FFN_dropout=args.FFN_dropout, FFN_layerNorm=args.FFN_layerNorm,
                         position_dropout=args.position_dropout, weight_init=args.encoder_weight_init,
                         position_encoding=args.position_encoding, CLS_learned=args.CLS_learned,
                         max_position_tokens=args.max_position_tokens, relative_pe=args.relative_position_encoding,
                         window_size=args.window_size, window_depth=args.part_len, conv_patch=args.conv_patch)   # :58-66
classifier_model = Classifier(args.d_model, args.classifier_dropout, weight_init=args.classifier_weight_init)   # :67
torch.save({"enc": temporal_model.state_dict(), "cls": classifier_model.state_dict()}, sys.argv[1] + ".init")
if args.data_parallel == True:                 # :76-78
    temporal_model = nn.DataParallel(temporal_model)
    classifier_model = nn.DataParallel(classifier_model)
temporal_model = temporal_model.cuda().train()  # :80-81
classifier_model = classifier_model.cuda().train()
optimizer = torch.optim.Adagrad([{"params": temporal_model.parameters(), "lr": args.lr_encoder},
                                 {"params": classifier_model.parameters(), "lr": args.lr_classifier}],
                                weight_decay=args.weight_decay)                                   # :83-85
g = torch.Generator().manual_seed(1)
log = []
for it in range(3):
    norm_feats = torch.randn(args.batch_size, args.part_num * args.part_len, args.n_patch, args.d_model, generator=g).abs()
    abnorm_feats = torch.randn(args.batch_size, args.part_num * args.part_len, args.n_patch, args.d_model, generator=g).abs()
    abnorm_labs = (torch.rand(args.batch_size, args.part_num * args.part_len, 1, generator=g) > 0.7).float()
    if it == 0:
        torch.save({"norm": norm_feats, "abn": abnorm_feats, "labs": abnorm_labs}, sys.argv[1] + ".batch0")
    # ---- :103-142, unmodified ----
    if args.temporal_only == False:
        norm_labs = torch.zeros([args.batch_size, args.part_num, 2], dtype=torch.float32).cuda()
        norm_labs[:, :, 0] += 1
        abnorm_labs = abnorm_labs.cuda().view([args.batch_size, args.part_num, args.part_len])
        abnorm_labs = torch.mean(abnorm_labs, dim=-1).view([args.batch_size, args.part_num, 1])
        abnorm_labs_tmp = torch.zeros([args.batch_size, args.part_num, 2], dtype=torch.float32).cuda()
        abnorm_labs_tmp[:, :, 1] = abnorm_labs[:, :, 0]
        abnorm_labs_tmp[:, :, 0] = 1 - abnorm_labs_tmp[:, :, 1]
        abnorm_labs = abnorm_labs_tmp
        clip_labs = torch.cat([norm_labs, abnorm_labs], dim=0)
    norm_feats = norm_feats.cuda().float().view([args.batch_size * args.part_num, args.part_len * args.n_patch, args.d_model])
    abnorm_feats = abnorm_feats.cuda().float().view([args.batch_size * args.part_num, args.part_len * args.n_patch, args.d_model])
    feats = torch.cat([norm_feats, abnorm_feats], dim=0)
    feats = temporal_model(feats)
    feats = feats[:, 0, :].float().view([args.batch_size * 2, args.part_num, args.d_model])
    outputs = classifier_model(feats)
    outputs = outputs.view([args.batch_size * 2 * args.part_num, -1])
    abnorm_score = outputs[:, 1]
    if args.temporal_only == False:
        clip_labs = clip_labs.view([args.batch_size * 2 * args.part_num, -1])
        CE_loss = get_CE_loss(args, outputs, clip_labs)
    else:
        CE_loss = 0
    MIL_loss, err, l1 = get_MIL_loss(args, abnorm_score)
    loss = args.lambda_MIL * MIL_loss + args.lambda_CE * CE_loss
    optimizer.zero_grad()
    loss.backward()
    if args.clip_grad == True:
        torch.nn.utils.clip_grad_norm_(temporal_model.parameters(), 10)
        torch.nn.utils.clip_grad_norm_(classifier_model.parameters(), 10)
    optimizer.step()
    log.append(dict(loss=float(loss), mil=float(MIL_loss), ce=float(CE_loss), l1=float(l1)))
sd = temporal_model.state_dict()
assert all(k.startswith("module.") for k in sd)          # DataParallel checkpoints carry the prefix the consumers strip
torch.save({k[7:]: v.cpu() for k, v in sd.items()}, sys.argv[1] + ".final")
print(json.dumps(log))
'''


def test_reference_train_loop_body_runs_on_the_dropin_modules(tmp_path):
    from oracle import lstc_oracle as O
    stem = str(tmp_path / "run")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin") + os.pathsep + ROOT)
    r = subprocess.run([sys.executable, "-c", LOOP, stem], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    log = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(log) == 3 and all(torch.isfinite(torch.tensor(list(x.values()))).all() for x in log)
    # first iteration against the CPU oracle on the same initial weights and batch
    init = torch.load(stem + ".init", weights_only=False)
    b0 = torch.load(stem + ".batch0", weights_only=False)
    B, P, T, N, D = 4, 4, 3, 16, 256
    cfg = O.EncoderConfig(n_layers=2, n_head=2, d_k=64, d_v=64, d_model=D, d_inner=512, MHA_layerNorm=True,
                          FFN_layerNorm=True, relative_pe=True, window_size=4, window_depth=T)
    feats = torch.cat([b0["norm"].view(B * P, T * N, D), b0["abn"].view(B * P, T * N, D)])
    labs = O.soft_labels(b0["labs"].view(B, P * T), B, P, T)
    ref_loss, aux = O.ltn_train_loss(init["enc"], init["cls"], feats, labs, cfg, B, P)
    assert abs(log[0]["loss"] - ref_loss.item()) < 1e-2, (log[0], ref_loss.item())
    assert abs(log[0]["mil"] - aux["mil"].item()) < 1e-2 and abs(log[0]["ce"] - aux["ce"].item()) < 1e-2
    # the stock optimizer moved the drop-in modules' parameters (weights AND the rel-pos bias table)
    final = torch.load(stem + ".final", weights_only=False)
    moved = {k: (final[k].float() - init["enc"][k].float()).abs().max().item() for k in final
             if final[k].is_floating_point()}
    assert moved["layer_stack.0.slf_attn.w_qs.weight"] > 0 and moved["layer_stack.1.pos_ffn.w_2.bias"] > 0
    assert moved["layer_stack.0.slf_attn.relative_position_bias_table"] > 0
    assert set(final) == set(init["enc"])
