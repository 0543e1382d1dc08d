"""Batched drivers around the drop-in modules: the LTN / STN train step (forward + losses + backward
[+ fused Adagrad]), data-parallel sharding by video pair with an overlapped bucketed gradient all-reduce,
and sharded, batched pseudo-label generation / scoring (no collective).

These replace the *loops* of the reference scripts (Train/temporal_transformer_shanghaitech.py:99-144 train
step, :163-226 and Train/pseudo_labels_generator_temporal.py:113-143 one-window-per-forward sweeps); the
scripts themselves keep working unmodified on lstc_vad_b200.models.
"""
from __future__ import annotations

import math
import os
import types
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import functional as Fn
from . import losses, ops
from .models import Classifier, Encoder, Regressor

# --------------------------------------------------------------------------------------------------
# workload configurations named by BASELINE.json (shapes: SURVEY.md §8 table)
# --------------------------------------------------------------------------------------------------
@dataclass
class Workload:
    name: str
    d_model: int
    d_inner: int
    part_len: int          # T clips per window
    n_patch: int           # N tokens per clip
    part_num: int          # P windows per video
    batch_size: int        # B video pairs per step
    relative_pe: bool = True
    window_size: int = 4
    MHA_layerNorm: bool = True
    FFN_layerNorm: bool = True
    kind: str = "ltn"      # "ltn" (Classifier, MIL + CE) or "stn" (Regressor, MIL)
    n_layers: int = 3
    n_head: int = 8
    d_k: int = 256
    dropouts: Tuple[float, float, float, float] = (0.2, 0.2, 0.1, 0.6)  # attn, fc, ffn, head

    @property
    def tokens_per_window(self) -> int:
        return (self.part_len if self.kind == "ltn" else 1) * self.n_patch

    @property
    def windows_per_step(self) -> int:
        w = 2 * self.batch_size * self.part_num
        return w if self.kind == "ltn" else w * self.part_len

    def encoder_kwargs(self) -> dict:
        a, f, n, _ = self.dropouts
        return dict(n_layers=self.n_layers, n_head=self.n_head, d_k=self.d_k, d_v=self.d_k, d_model=self.d_model,
                    d_inner=self.d_inner, MHA_attn_dropout=a, MHA_fc_dropout=f, MHA_layerNorm=self.MHA_layerNorm,
                    FFN_dropout=n, FFN_layerNorm=self.FFN_layerNorm, weight_init=(self.kind == "stn"),
                    relative_pe=self.relative_pe and self.kind == "ltn", window_size=self.window_size,
                    window_depth=self.part_len)

    def fwd_flops_per_window(self) -> float:
        """Algorithmic forward FLOPs per window: per layer 2L*D*3*HD + 4*H*L^2*dk + 2L*HD*D +
        4L*D*Dh, plus the head 2(512 D + 16384 + 32 C)."""
        L = self.tokens_per_window + 1
        D, Dh, H, dk = self.d_model, self.d_inner, self.n_head, self.d_k
        HD = H * dk
        layer = 2 * L * D * 3 * HD + 4 * H * L * L * dk + 2 * L * HD * D + 4 * L * D * Dh
        return self.n_layers * layer + 2 * (512 * D + 512 * 32 + 32 * 2)


WORKLOADS: Dict[str, Workload] = {
    # configs[1] of BASELINE.json: LTN train step at the ShanghaiTech shape — the headline configuration
    "ltn_sht": Workload("ltn_sht", 2048, 4096, 3, 16, 16, 40),
    "ltn_ucf": Workload("ltn_ucf", 2048, 4096, 2, 9, 32, 40),
    "ltn_ubnormal": Workload("ltn_ubnormal", 1024, 4096, 5, 16, 16, 40),
    "stn_sht": Workload("stn_sht", 2048, 3027, 7, 16, 16, 40, relative_pe=False, MHA_layerNorm=False, kind="stn",
                        dropouts=(0.1, 0.1, 0.3, 0.6)),
}


def synthetic_step_inputs(wl: Workload, seed: int = 0, batch_size: Optional[int] = None, device="cpu",
                          pin: bool = False):
    """Non-negative heavy-tailed features (I3D features are post-ReLU) + per-clip pseudo labels.
    Returns (feats fp32 [W, L0, D] normal windows first, soft labels fp32 [W, 2] | None)."""
    B = batch_size or wl.batch_size
    P, T = wl.part_num, wl.part_len
    W = 2 * B * P * (1 if wl.kind == "ltn" else T)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(W, wl.tokens_per_window, wl.d_model, generator=g).abs_()
    labs = None
    if wl.kind == "ltn":
        pseudo = (torch.rand(B, P * T, generator=g) > 0.9).float() * torch.rand(B, P * T, generator=g)
        labs = losses.soft_clip_labels(pseudo, B, P, T)
    if pin:
        feats = feats.pin_memory()
        labs = labs.pin_memory() if labs is not None else None
    if device != "cpu":
        feats = feats.to(device, non_blocking=True)
        labs = labs.to(device, non_blocking=True) if labs is not None else None
    return feats, labs


# --------------------------------------------------------------------------------------------------
# single-GPU / per-rank train step
# --------------------------------------------------------------------------------------------------
class TrainStep:
    """Forward + loss + backward of one reference train step on this rank's share of the batch.

    With `world_size > 1` the batch is sharded by VIDEO PAIR (all P windows of a bag stay on one rank); the
    MIL hinge couples bags across ranks, so the per-window scores (a few KB) are all-gathered and every rank
    evaluates the global MIL loss and keeps the gradient slice of its own windows; CE is a share of the global
    mean.  Parameter gradients are then SUM-all-reduced in buckets, overlapped with the rest of backward."""

    def __init__(self, wl: Workload, device: torch.device, seed: int = 0, train_mode: bool = True,
                 lambda_1: float = 0.01, lambda_MIL: float = 1.0, lambda_CE: float = 0.8,
                 process_group=None, optimizer: bool = False, lr_encoder: float = 1e-4, lr_head: float = 1e-2,
                 weight_decay: float = 1e-3, clip_grad_norm: Optional[float] = None, cls_fast_path: bool = False,
                 dp_grad_dtype: Optional[str] = None):
        self.wl, self.device = wl, device
        self.lambda_1, self.lambda_MIL, self.lambda_CE = lambda_1, lambda_MIL, lambda_CE
        self.cls_fast_path = cls_fast_path  # Encoder.forward_cls: skips the last layer's dead (non-CLS) work
        self.pg = process_group
        self.world = 1
        self.rank = 0
        if process_group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        torch.manual_seed(seed)  # identical initial weights on every rank
        self.encoder = Encoder(**wl.encoder_kwargs()).to(device)
        head_drop = wl.dropouts[3]
        self.head = (Classifier(wl.d_model, head_drop) if wl.kind == "ltn" else Regressor(wl.d_model, head_drop)).to(device)
        self.encoder.train(train_mode)
        self.head.train(train_mode)
        Fn.set_dropout_stream(seed * 1000003 + 7919 * self.rank + 1)
        self.reducer = None
        if self.world > 1:
            self.reducer = BucketedGradReducer([self.encoder, self.head], process_group, grad_dtype=dp_grad_dtype)
        self.opt = None
        if optimizer:
            self.opt = FusedAdagrad([(list(self.encoder.parameters()), lr_encoder),
                                     (list(self.head.parameters()), lr_head)], weight_decay,
                                    clip_grad_norm=clip_grad_norm)

    def parameters(self):
        return list(self.encoder.parameters()) + list(self.head.parameters())

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    def forward_backward(self, feats: torch.Tensor, labs: Optional[torch.Tensor], local_batch: int) -> Dict[str, torch.Tensor]:
        """feats [W_local, L0, D] on device (normal windows of this rank's bags first, then abnormal);
        labs [W_local, 2].  Returns the (global) loss terms as device scalars."""
        wl = self.wl
        B, P, T = local_batch, wl.part_num, wl.part_len
        args = types.SimpleNamespace(batch_size=B, part_num=P, part_len=T, lambda_1=self.lambda_1)
        if self.reducer is not None:
            self.reducer.start()
        if self.cls_fast_path:
            cls_rows = self.encoder.forward_cls(feats)
        else:
            out = self.encoder(feats)
            cls_rows = out[:, 0, :]
        if wl.kind == "ltn":
            probs = self.head(cls_rows.view(2 * B, P, wl.d_model)).view(2 * B * P, -1)
            score = probs[:, 1]
            if self.world == 1:
                ce = losses.get_CE_loss(args, probs, labs)
                mil, err, l1 = losses.get_MIL_loss(args, score)
            else:
                ce = losses.get_CE_loss(args, probs, labs) / self.world
                mil, err, l1 = self._global_mil(score, B, P, 1, flat=True)
            loss = self.lambda_MIL * mil + self.lambda_CE * ce
            terms = dict(loss=loss, mil=mil, ce=ce, err=err, spar=l1, scores=score)
        else:
            score = self.head(cls_rows.view(2 * B, P * T, wl.d_model)).view(2 * B, P * T, 1)
            if self.world == 1:
                mil, err, l1 = losses.get_MIL_loss(args, score)
            else:
                mil, err, l1 = self._global_mil(score, B, P, T, flat=False)
            loss = mil
            terms = dict(loss=loss, mil=mil, err=err, spar=l1, scores=score)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        if self.opt is not None:
            self.opt.step()
        return terms

    # -- MIL across ranks: all-gather the scores, evaluate the global loss everywhere, keep the local slice --
    def _global_mil(self, score: torch.Tensor, B: int, P: int, T: int, flat: bool):
        import torch.distributed as dist
        flat_local = score.reshape(-1).contiguous()
        gathered = [torch.empty_like(flat_local) for _ in range(self.world)]
        dist.all_gather(gathered, flat_local.detach(), group=self.pg)
        glob = global_score_order(gathered)
        Bg = B * self.world
        spar_start = Bg if flat else Bg * P * T
        out3, _, dglob = ops.mil_loss(glob, Bg, P, T, 1, self.lambda_1, spar_start)
        d_local = local_grad_slice(dglob, self.rank, self.world)
        mil = _InjectGrad.apply(flat_local, out3[0], d_local)
        return mil, out3[1], out3[2]


def global_score_order(gathered: Sequence[torch.Tensor]) -> torch.Tensor:
    """Per-rank score vectors (each: this rank's normal windows then its abnormal windows) -> the reference's flat
    order for the whole batch: every normal bag (rank-major) first, then every abnormal bag."""
    half = gathered[0].numel() // 2
    return torch.cat([g[:half] for g in gathered] + [g[half:] for g in gathered])


def local_grad_slice(dglob: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Inverse of global_score_order for one rank: the gradient entries of this rank's windows, local order."""
    half = dglob.numel() // (2 * world)
    n_norm = half * world
    return torch.cat([dglob[rank * half:(rank + 1) * half], dglob[n_norm + rank * half:n_norm + (rank + 1) * half]])


class GraphedTrainStep:
    """One CUDA graph for the whole train step (forward, losses, backward and, if the TrainStep owns one, the fused
    optimizer step) at fixed shapes: ~110 kernel launches and the autograd bookkeeping are replayed by the driver
    instead of being re-issued from Python.

    Dropout: a graph bakes every kernel's (seed, offset) by value, so a device-resident step counter is registered
    with the library (`lstc_set_rng_step`) and bumped INSIDE the graph; every replay therefore draws fresh masks.
    The bf16 weight copies are re-cast inside the graph as well (the cache is invalidated right before capture), so
    an in-graph optimizer step is seen by the next replay.  With world_size > 1 the score all-gather and the bucketed
    side-stream all-reduces are captured too (NCCL collectives are graph-capturable); every rank must construct and
    replay its graph in lock-step, and the process should exit without destroying the process group while a graph
    that holds captured collectives is alive."""

    RNG_STRIDE = 4096  # > number of dropout sites per step

    def __init__(self, step: "TrainStep", feats: torch.Tensor, labs: Optional[torch.Tensor], local_batch: int,
                 warmup: int = 3):
        # Autograd's AccumulateGrad nodes remember the stream of the forward that created them and live as long as any
        # autograd graph of an earlier (eager, default-stream) step is referenced; such a node would make the legacy
        # stream wait on the capturing stream.  Drop dead graphs so the side-stream warm-up below re-creates the nodes.
        # (Callers must not keep loss tensors of earlier eager steps alive across this constructor.)
        import gc
        step.zero_grad()
        gc.collect()
        self.step, self.local_batch = step, local_batch
        self.static_feats = feats.clone()
        self.static_labs = labs.clone() if labs is not None else None
        # process-wide per-device counter (shared by all graphs; see functional.device_rng_counter)
        self.rng_counter = Fn.device_rng_counter(feats.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step.zero_grad()
                step.forward_backward(self.static_feats, self.static_labs, local_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        Fn.invalidate_weight_cache()
        step.zero_grad()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if step.world > 1 else "global"):
            self.rng_counter += self.RNG_STRIDE
            self.terms = step.forward_backward(self.static_feats, self.static_labs, local_batch)

    def __call__(self, feats: Optional[torch.Tensor] = None, labs: Optional[torch.Tensor] = None):
        """Replays the step; new inputs (same shapes) are copied into the graph's static buffers first.  The returned
        loss terms and the parameters' .grad tensors are the graph's static outputs."""
        if feats is not None:
            self.static_feats.copy_(feats, non_blocking=True)
        if labs is not None and self.static_labs is not None:
            self.static_labs.copy_(labs, non_blocking=True)
        self.graph.replay()
        return self.terms

    def close(self):
        """Releases the captured graph (its collectives, static buffers and autograd state).  The dropout step counter
        is shared by every graph of the process and stays registered."""
        self.graph = None
        self.terms = None


class _InjectGrad(torch.autograd.Function):
    """value (precomputed global loss) whose gradient w.r.t. `scores` is the given slice."""

    @staticmethod
    def forward(ctx, scores, value, dscores):
        ctx.save_for_backward(dscores)
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return ops.scale_by_device_scalar(d, g.reshape(1).contiguous()), None, None


# --------------------------------------------------------------------------------------------------
# bucketed gradient all-reduce overlapped with backward (replaces nn.DataParallel's reduce_add_coalesced,
# Train/temporal_transformer_shanghaitech.py:76-78)
# --------------------------------------------------------------------------------------------------
class BucketedGradReducer:
    """SUM all-reduce of the parameter gradients over the data-parallel ranks through flat bf16 (or fp32) buckets: one
    multi-tensor pack launch, one NCCL all-reduce, one unpack launch per bucket.  Default plan: a single bucket reduced
    after backward; LSTC_DP_BUCKETS=layer gives two buckets per EncoderLayer (FFN, attention) + one for everything
    else, which LSTC_DP_REDUCE=overlap launches in reverse layer order from the backward hooks.  The plan is static and
    only contains parameters that receive gradients (LayerNorms switched off by the model flags never do - SURVEY §8)."""

    def __init__(self, modules: Sequence[torch.nn.Module], process_group, comm_stream: Optional[torch.cuda.Stream] = None,
                 grad_dtype: Optional[str] = None):
        import torch.distributed as dist
        self.dist, self.pg = dist, process_group
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._plan(modules)
        self.flat: List[Optional[tuple]] = [None] * len(self.buckets)  # (param ids, flat buffer, offsets)
        # payload dtype of the gradient all-reduce: "bf16" (default; fp32 gradients are rounded once before the sum,
        # relative error ~2^-9 per element, well inside the bf16 activation noise of the step) or "fp32" (exact sum,
        # as nn.DataParallel's reduce_add_coalesced)
        self.grad_dtype = grad_dtype or os.environ.get("LSTC_DP_GRAD_DTYPE", "bf16")
        if self.grad_dtype not in ("bf16", "fp32"):
            raise ValueError(f"LSTC_DP_GRAD_DTYPE must be bf16 or fp32, got {self.grad_dtype!r}")
        self.pending: List[int] = [0] * len(self.buckets)
        self.handles = []
        self.stream = comm_stream or torch.cuda.Stream()
        self._active = False
        self._skip: set = set()
        # "tail" (default): launch the bucket(s) after backward (no SM contention between NCCL and the persistent GEMMs);
        # "overlap": launch a bucket's all-reduce from the backward hooks as soon as its last gradient lands
        self.mode = os.environ.get("LSTC_DP_REDUCE", "tail")
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                p.register_post_accumulate_grad_hook(self._make_hook(bi))

    def _plan(self, modules):
        if os.environ.get("LSTC_DP_BUCKETS", "single") == "single":
            # default: ONE flat bucket for the whole model (203 MB of bf16 for the LTN), reduced right after backward
            # (LSTC_DP_REDUCE=tail): a single large all-reduce runs at NVLink / NVSwitch bus bandwidth, where several
            # ~30 MB ones pay launch latency each, and NCCL's CTAs do not fight the persistent one-CTA-per-SM GEMMs of
            # backward for SMs.  Measured on 8 B200 (bench.py, CUDA graph, 20 steps): 30.70 ms / step against 31.07 ms
            # for per-layer buckets launched from the backward hooks (LSTC_DP_BUCKETS=layer LSTC_DP_REDUCE=overlap)
            # and 31.11 ms for per-layer buckets launched after backward.
            ps = [p for m in modules for p in m.parameters() if p.requires_grad]
            if ps:
                self.buckets.append(ps)
            return
        rest = []
        for m in modules:
            stack = getattr(m, "layer_stack", None)
            in_layers = set()
            if stack is not None:
                for layer in stack:
                    # two buckets per layer: the FFN gradients are complete before the attention block's backward
                    # starts, so their all-reduce overlaps it; the last exposed bucket is half a layer, not a layer
                    subs = [getattr(layer, "pos_ffn", None), getattr(layer, "slf_attn", None)]
                    seen = set()
                    for sub in subs:
                        if sub is None:
                            continue
                        ps = [p for p in sub.parameters() if p.requires_grad]
                        seen.update(id(p) for p in ps)
                        if ps:
                            self.buckets.append(ps)
                    rest_l = [p for p in layer.parameters() if p.requires_grad and id(p) not in seen]
                    if rest_l:
                        self.buckets.append(rest_l)
                    in_layers.update(id(p) for p in layer.parameters())
            rest += [p for p in m.parameters() if p.requires_grad and id(p) not in in_layers]
        if rest:
            self.buckets.append(rest)

    def mark_unused(self, params):
        """Parameters that never receive a gradient (found by a warm-up step) are dropped from the plan."""
        self._skip.update(id(p) for p in params)

    def _make_hook(self, bi):
        def hook(p):
            if not self._active:
                return
            self.pending[bi] -= 1
            if self.pending[bi] == 0 and self.mode != "tail":
                self._launch(bi)
        return hook

    def start(self):
        self._active = True
        self.handles = []
        for bi, bucket in enumerate(self.buckets):
            self.pending[bi] = sum(1 for p in bucket if id(p) not in self._skip)

    ALIGN = 64  # elements: every tensor starts on a 128-byte (bf16) / 256-byte (fp32) boundary of its flat bucket

    def _layout(self, bi, ps):
        """(flat buffer, per-tensor offsets) of bucket `bi` for the parameter list `ps` (static after the first steps)."""
        key = tuple(id(p) for p in ps)
        cached = self.flat[bi]
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        offsets, off = [], 0
        for p in ps:
            offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dtype = torch.bfloat16 if self.grad_dtype == "bf16" else torch.float32
        flat = torch.zeros(off, device=ps[0].device, dtype=dtype)  # zeros: the alignment gaps travel through the reduce
        self.flat[bi] = (key, flat, offsets)
        return flat, offsets

    def _launch(self, bi):
        ps = [p for p in self.buckets[bi] if id(p) not in self._skip and p.grad is not None]
        if not ps:
            return
        flat, offsets = self._layout(bi, ps)
        grads = [p.grad for p in ps]
        views = None
        if self.grad_dtype == "bf16":
            # one launch: fp32 gradients -> the flat bf16 all-reduce payload (half the NVLink bytes of an fp32 bucket)
            ops.multi_pack_bf16(grads, flat, offsets)
        else:
            views = [flat[o:o + p.numel()].view_as(p) for o, p in zip(offsets, ps)]
            torch._foreach_copy_(views, grads)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.pg)
            # the reduced gradient goes back into the parameters' own .grad tensors (in place, on the side stream, so
            # it overlaps the rest of backward as well): no tensor is re-bound, the optimizer sees ordinary fp32 grads
            if self.grad_dtype == "bf16":
                ops.multi_unpack_bf16(flat, offsets, grads)
            else:
                torch._foreach_copy_(grads, views)
        self.handles.append(bi)

    def finish(self):
        # buckets whose parameters did not all fire (unused params): flush them now
        for bi in range(len(self.buckets)):
            if self.pending[bi] > 0:
                unused = [p for p in self.buckets[bi] if p.grad is None]
                self.mark_unused(unused)
                self.pending[bi] = 0
                self._launch(bi)
            elif self.mode == "tail":
                self._launch(bi)
        torch.cuda.current_stream().wait_stream(self.stream)
        self._active = False


class FusedAdagrad:
    """torch.optim.Adagrad(lr per group, weight_decay) semantics (Train/temporal_transformer_shanghaitech.py:83-85) as
    ONE multi-tensor launch over every parameter that has a gradient (5 HBM passes: read grad / param / state, write
    param / state); with `clip_grad_norm` one launch per parameter group, each consuming its own device-side clip
    coefficient (the scripts clip per model, :139-141)."""

    def __init__(self, groups, weight_decay: float, eps: float = 1e-10, clip_grad_norm: Optional[float] = None):
        self.groups = groups
        self.wd, self.eps = weight_decay, eps
        self.clip = clip_grad_norm  # per group, like the scripts' per-model clip_grad_norm_(.., 10)
        self.state: Dict[int, torch.Tensor] = {}

    def _state(self, p):
        st = self.state.get(id(p))
        if st is None:
            st = self.state[id(p)] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    def step(self):
        ps, gs, sts, lrs = [], [], [], []
        for params, lr in self.groups:
            live = [p for p in params if p.grad is not None]
            if not live:
                continue
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in live]
            states = [self._state(p) for p in live]
            if self.clip:
                coef = ops.grad_clip_coef(grads, self.clip)
                ops.multi_adagrad([p.data for p in live], grads, states, [lr] * len(live), self.wd, self.eps, 1.0, coef)
            else:
                ps += [p.data for p in live]
                gs += grads
                sts += states
                lrs += [lr] * len(live)
        if ps:
            ops.multi_adagrad(ps, gs, sts, lrs, self.wd, self.eps)
        Fn.invalidate_weight_cache()  # parameters were updated through raw pointers


# --------------------------------------------------------------------------------------------------
# batched, sharded scoring / pseudo-label generation (no collective)
# --------------------------------------------------------------------------------------------------
def video_windows(n_clips: int, part_len: int, backshift: bool) -> List[Tuple[int, int, int]]:
    """(feat_beg, feat_end, n_clips_covered) per window.  backshift=False: short trailing window
    (Train/pseudo_labels_generator_temporal.py:127-134); True: trailing window moved back to full length while
    its score still covers only the remaining clips (Test/evaluation_shanghaitech_ubnormal.py:74-92)."""
    out = []
    n_win = (n_clips + part_len - 1) // part_len
    for i in range(n_win):
        beg = i * part_len
        end = n_clips if i == n_win - 1 else (i + 1) * part_len
        if backshift and end - beg < part_len and end - part_len >= 0:
            out.append((end - part_len, end, end - beg))
        else:
            out.append((beg, end, end - beg))
    return out


def window_plan(n_clips: int, part_len: int, backshift: bool) -> Tuple[int, Optional[Tuple[int, int, int]]]:
    """The same windows as `video_windows`, in the form the batched scorer wants: (number of full windows, which are
    the contiguous clips [0, n_full*part_len) and can be taken as ONE strided view, and the ragged trailing window
    (feat_beg, feat_end, n_clips_covered) or None)."""
    n_full, rem = divmod(n_clips, part_len)
    if rem == 0:
        return n_full, None
    if backshift and n_clips >= part_len:
        return n_full, (n_clips - part_len, n_clips, rem)
    return n_full, (n_full * part_len, n_clips, rem)


def shard_videos(keys: Sequence[str], n_clips: Sequence[int], world: int, rank: int) -> List[int]:
    """Greedy longest-first assignment of videos to ranks (balanced by clip count); returns this rank's indices."""
    order = sorted(range(len(keys)), key=lambda i: (-n_clips[i], keys[i]))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        load[r] += n_clips[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


@torch.no_grad()
def score_videos(encoder: Encoder, head: torch.nn.Module, videos: Dict[str, torch.Tensor], part_len: int,
                 backshift: bool = False, threshold: Optional[float] = None, max_windows: int = 4096,
                 n_patch: Optional[int] = None, cls_fast_path: bool = False) -> Dict[str, torch.Tensor]:
    """Per-clip scores for every video of this rank's shard: {key: fp32 [n_clips]} on the CPU.

    Windows of equal token count from ALL videos are batched into a few large forward calls instead of the
    reference's one-window-per-forward loop; `threshold` applies where(score > thr, score, 0)."""
    encoder.eval()
    head.eval()
    dev = next(encoder.parameters()).device
    is_cls = isinstance(head, Classifier)
    T = part_len
    # windows grouped by clip count: all full windows of a video are ONE view [n_full, T*N, D] of its feature tensor
    # (no per-window slicing on the host); the ragged trailing window (short, or back-shifted to full length) joins
    # the group of its own length
    groups: Dict[int, List[Tuple[str, int, torch.Tensor]]] = {}   # clips per window -> [(key, first window index, [n, c*N, D])]
    for key, feats in videos.items():
        f = feats[:, :n_patch] if n_patch is not None else feats
        D = f.shape[-1]
        n_full, tail = window_plan(f.shape[0], T, backshift)
        if n_full:
            groups.setdefault(T, []).append((key, 0, f[:n_full * T].reshape(n_full, -1, D)))
        if tail is not None:
            beg, end, _ = tail
            groups.setdefault(end - beg, []).append((key, n_full, f[beg:end].reshape(1, -1, D)))
    win_scores: Dict[str, Dict[int, torch.Tensor]] = {}
    def run(xin):
        sc = head(encoder.forward_cls(xin) if cls_fast_path else encoder(xin)[:, 0, :])
        sc = sc[:, 1] if is_cls else sc[:, 0]
        if threshold is not None:
            sc = losses.threshold_pseudo_labels(sc, threshold)
        return sc.float()

    for _, items in groups.items():
        # forward calls of up to `max_windows` windows, assembled from consecutive videos: the concatenated copy of the
        # inputs never exceeds one call's worth (a corpus-sized torch.cat would double the resident features)
        outs, pending, n_pending = [], [], 0

        def flush():
            nonlocal pending, n_pending
            if pending:
                xin = (pending[0] if len(pending) == 1 else torch.cat(pending)).to(dev, non_blocking=True).float()
                outs.append(run(xin))
                pending, n_pending = [], 0

        for _, _, x in items:
            s0 = 0
            while s0 < x.shape[0]:
                take = min(x.shape[0] - s0, max_windows - n_pending)
                pending.append(x[s0:s0 + take])
                n_pending += take
                s0 += take
                if n_pending == max_windows:
                    flush()
        flush()
        sc = torch.cat(outs).cpu()
        off = 0
        for key, w0, x in items:
            win_scores.setdefault(key, {})[w0] = sc[off:off + x.shape[0]]
            off += x.shape[0]
    result = {}
    for key, feats in videos.items():
        n = feats.shape[0]
        n_full, rem = n // T, n % T
        parts = []
        if n_full:
            parts.append(win_scores[key][0].repeat_interleave(T))
        if rem:
            parts.append(win_scores[key][n_full].repeat_interleave(rem))
        result[key] = torch.cat(parts).to(torch.float32) if parts else torch.zeros(0)
    return result


def save_pseudo_labels(path: str, scores: Dict[str, torch.Tensor]) -> None:
    """Writes {key + '.npy': float32 [n_clips, 1]} with np.save — the pickled-dict format the reference's generators
    produce (Train/pseudo_labels_generator_temporal.py:142-145) and its datasets read back
    (utils/load_dataset.py:17-25)."""
    import numpy as np
    out = {k + ".npy": v.detach().cpu().numpy().astype(np.float32).reshape(-1, 1) for k, v in scores.items()}
    np.save(path, out)


def frame_scores(clip_scores: torch.Tensor, segment_len: int = 16) -> torch.Tensor:
    """Per-clip scores -> per-frame scores (each clip covers `segment_len` frames), as the evaluation loop does
    (Test/evaluation_shanghaitech_ubnormal.py:92)."""
    return clip_scores.repeat_interleave(segment_len)


def pool_video_bins(feats: torch.Tensor, n_bins: int = 32, l2norm: bool = True) -> Tuple[torch.Tensor, List[int]]:
    """UCF-Crime evaluation pooling (Test/evaluation_UCF.py:52-77): a video of n_clips clips becomes `n_bins` pseudo
    clips, bin b = mean of clips [r[b], r[b+1]) with r = linspace(0, n_clips, n_bins+1) truncated to int32 (the single
    clip r[b] when the bin is empty), every token L2-normalised.  Returns (pooled fp32 [n_bins, n_patch, D] on the
    GPU, r as a list — the evaluation loop needs it to expand bin scores back to frames)."""
    import numpy as np
    n_clips = feats.shape[0]
    r = np.linspace(0, n_clips, n_bins + 1, dtype=np.int32)
    bounds = torch.from_numpy(r).to(feats.device)
    return ops.segment_mean(feats.float(), bounds, l2norm), r.tolist()


def frame_level_auc(window_scores: torch.Tensor, pos_frames: torch.Tensor, neg_frames: torch.Tensor) -> float:
    """Frame-level ROC-AUC of a scored corpus without materialising per-frame arrays or leaving the GPU for the sort:
    window i has score window_scores[i] and covers pos_frames[i] anomalous + neg_frames[i] normal frames (the reference
    repeats each window score over its frames and calls sklearn, Test/evaluation_shanghaitech_ubnormal.py:92-96,
    utils/eval_utils.py:21-24).  Up to 16384 windows per call."""
    dev = window_scores.device
    out = ops.weighted_auc(window_scores.float(), pos_frames.to(dev).float(), neg_frames.to(dev).float())
    return float(out[0].item())


# --------------------------------------------------------------------------------------------------
# training-window sampler on a device-resident corpus (utils/load_dataset.py:56-88, :90-108)
# --------------------------------------------------------------------------------------------------
def sample_window_indices(feat_len: int, part_num: int, part_len: int, sample: str = "uniform", rng=None):
    """Clip indices of the `part_num` windows of `part_len` consecutive clips the reference draws from a video of
    `feat_len` clips — same numpy calls in the same order as `sample_feat` (utils/load_dataset.py:68-88), so a seeded
    `np.random` / RandomState reproduces the reference's choice.  Returns int64 [part_num * part_len]."""
    import numpy as np
    rng = rng if rng is not None else np.random
    if feat_len < part_len:
        # the reference's numpy fancy-indexing raises (or silently wraps negative indices) here; a device gather would
        # read out of bounds instead, so refuse up front
        raise IndexError(f"sample_window_indices: video of {feat_len} clips is shorter than part_len {part_len}")
    base = np.linspace(0, feat_len - part_len, num=part_num + 1, dtype=int)
    if sample == "uniform":
        span = (feat_len - part_len) // (part_num + 1)
        move = 0 if span < 1 else rng.randint(span)
        chosen = base + move
        chosen = chosen.repeat(part_len).reshape([-1, part_len]) + np.arange(0, part_len, 1, dtype=int)
    else:
        chosen = base.repeat(part_len).reshape([-1, part_len]) + np.arange(0, part_len, 1, dtype=int)
        gap = chosen[1, 0] - chosen[0, 0]
        if gap == 0:
            move = 0
        else:
            move = rng.randint(0, gap, [part_num + 1]).repeat(part_len).reshape([-1, part_len])
        chosen = chosen + move
    chosen = chosen.reshape([-1])[: part_num * part_len].astype("int64")
    if chosen.min() < 0 or chosen.max() >= feat_len:
        raise IndexError(f"sample_window_indices: clip index out of range [0, {feat_len}) "
                         f"(min {chosen.min()}, max {chosen.max()})")
    return chosen


class DeviceCorpus:
    """Pre-extracted features of a training set kept resident in HBM (the ShanghaiTech train set is 238 videos; the
    reference holds them in host RAM and ships 0.5 GB over PCIe every step, utils/load_dataset.py:29-48,
    Train/temporal_transformer_shanghaitech.py:115-118).  `sample_batch` draws B normal + B abnormal videos, picks their
    windows like the reference's `sample_feat` and gathers them on the device into the [2*B*P, T*N, D] step input."""

    def __init__(self, normal: Dict[str, torch.Tensor], abnormal: Dict[str, torch.Tensor], device,
                 pseudo_labels: Optional[Dict[str, torch.Tensor]] = None):
        self.device = device
        self.norm_keys, self.abn_keys = sorted(normal), sorted(abnormal)
        self.norm = [normal[k].to(device).float().contiguous() for k in self.norm_keys]
        self.abn = [abnormal[k].to(device).float().contiguous() for k in self.abn_keys]
        self.pseudo = pseudo_labels

    def __len__(self):
        return min(len(self.norm), len(self.abn))

    def sample_batch(self, batch_size: int, part_num: int, part_len: int, sample: str = "uniform", rng=None):
        import numpy as np
        rng = rng if rng is not None else np.random
        ni = rng.permutation(len(self.norm))[:batch_size]
        ai = rng.permutation(len(self.abn))[:batch_size]
        feats, abn_labels = [], []
        for which, ids in (("n", ni), ("a", ai)):
            for i in ids:
                f = self.norm[i] if which == "n" else self.abn[i]
                idx = sample_window_indices(f.shape[0], part_num, part_len, sample, rng)
                sel = ops.gather_rows(f, torch.from_numpy(idx).to(self.device))       # [P*T, N, D]
                feats.append(sel.view(part_num, part_len * f.shape[1], f.shape[2]))
                if which == "a":
                    key = self.abn_keys[i]
                    if self.pseudo is not None and key in self.pseudo:
                        lab = self.pseudo[key].reshape(-1).float()[torch.from_numpy(idx)]
                    else:
                        lab = torch.ones(part_num * part_len)
                    abn_labels.append(lab)
        feats = torch.cat(feats, dim=0)
        labs = losses.soft_clip_labels(torch.stack(abn_labels), batch_size, part_num, part_len).to(self.device)
        return feats, labs


# --------------------------------------------------------------------------------------------------
# one co-teaching round (README.md:21-36 of the reference: the four scripts run back to back)
# --------------------------------------------------------------------------------------------------
def synthetic_corpus(n_videos: int, n_patch: int, d_model: int, device, seed: int = 0, mean_clips: float = 50.0,
                     min_clips: int = 24, max_clips: int = 160):
    """(normal, abnormal) dicts {key: fp32 [n_clips, n_patch, d_model]} generated directly in HBM: half the videos
    abnormal, clip counts log-normal around `mean_clips`, non-negative heavy-tailed
    features like post-ReLU I3D activations."""
    import numpy as np
    rs = np.random.RandomState(seed)
    g = torch.Generator(device=device).manual_seed(seed)
    normal, abnormal = {}, {}
    for i in range(n_videos):
        n = int(np.clip(rs.lognormal(np.log(mean_clips), 0.5), min_clips, max_clips))
        f = torch.randn(n, n_patch, d_model, device=device, generator=g).abs_()
        (abnormal if i % 2 else normal)[f"v{i:05d}"] = f
    return normal, abnormal


def synthetic_video_lengths(n_videos: int, seed: int = 0, mean_clips: float = 50.0, min_clips: int = 24,
                            max_clips: int = 160) -> List[int]:
    """Clip counts of the synthetic corpus of SURVEY.md §8d (config C5): log-normal around `mean_clips`.  Host-only, so
    every rank can plan the sharding of the whole corpus without generating it."""
    import numpy as np
    rs = np.random.RandomState(seed)
    return [int(np.clip(rs.lognormal(np.log(mean_clips), 0.5), min_clips, max_clips)) for _ in range(n_videos)]


def synthetic_video(index: int, n_clips: int, n_patch: int, d_model: int, device, seed: int = 0) -> torch.Tensor:
    """Features of video `index` of the synthetic corpus, fp32 [n_clips, n_patch, d_model], generated in HBM from a
    per-video seed: the same tensor whatever rank (and however many ranks) generates it."""
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 7 * index + 1)
    return torch.randn(n_clips, n_patch, d_model, device=device, generator=g).abs_()


def sharded_label_sweep(encoder: Encoder, head: torch.nn.Module, n_videos: int, part_len: int, n_patch: int,
                        world: int = 1, rank: int = 0, threshold: Optional[float] = 0.65, seed: int = 0,
                        backshift: bool = False, cls_fast_path: bool = False, max_windows: int = 4096,
                        only_abnormal: bool = False, videos: Optional[Dict[str, torch.Tensor]] = None):
    """This rank's share of the pseudo-label generation sweep over the synthetic corpus (the loop of
    Train/pseudo_labels_generator_temporal.py:113-143 over every training video, batched and sharded by video with NO
    collective): the corpus is planned on the host (`synthetic_video_lengths`), assigned with `shard_videos` (greedy,
    balanced by clip count) and only this rank's videos are generated and scored.
    Returns (scores {key: fp32 [n_clips]}, stats dict with the device seconds of the sweep and its window count)."""
    d_model = encoder.layer_norm.weight.shape[0]
    dev = next(encoder.parameters()).device
    lengths = synthetic_video_lengths(n_videos, seed)
    keys = [f"v{i:05d}" for i in range(n_videos)]
    ids = [i for i in range(n_videos) if (i % 2 == 1 or not only_abnormal)]
    mine = shard_videos([keys[i] for i in ids], [lengths[i] for i in ids], world, rank)
    if videos is None:
        videos = {keys[ids[j]]: synthetic_video(ids[j], lengths[ids[j]], n_patch, d_model, dev, seed) for j in mine}
    n_windows = sum(-(-int(v.shape[0]) // part_len) for v in videos.values())
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scores = score_videos(encoder, head, videos, part_len=part_len, threshold=threshold, n_patch=n_patch,
                          backshift=backshift, cls_fast_path=cls_fast_path, max_windows=max_windows)
    e1.record()
    torch.cuda.synchronize(dev)
    return scores, dict(seconds=e0.elapsed_time(e1) * 1e-3, windows=n_windows, videos=len(videos),
                        clips=sum(int(v.shape[0]) for v in videos.values()))


def merge_label_dicts(local: Dict[str, torch.Tensor], process_group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Rank 0 receives every rank's {key: scores} dict and returns their union in key order (the other ranks return
    None) - the only exchange of the sharded sweep, after the timed region
    (Train/pseudo_labels_generator_temporal.py:145 then saves ONE dict)."""
    if process_group is None:
        return dict(sorted(local.items()))
    import torch.distributed as dist
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object({k: v.cpu() for k, v in local.items()}, gathered, dst=0, group=process_group)
    if rank != 0:
        return None
    merged: Dict[str, torch.Tensor] = {}
    for d in gathered:
        for k, v in d.items():
            if k in merged:
                raise RuntimeError(f"merge_label_dicts: video {k} was scored by two ranks")
            merged[k] = v
    return dict(sorted(merged.items()))


def co_teaching_round(corpus: "DeviceCorpus", stn: "TrainStep", ltn: "TrainStep", steps_per_epoch: Optional[int] = None,
                      thr_stn: float = 0.9, thr_ltn: float = 0.65, rng=None, cls_fast_path: bool = False) -> Dict:
    """STN epoch (MIL, Train/spatio_transformer_shanghaitech.py) -> STN pseudo labels, one clip per window, thr 0.9
    (Train/pseudo_labels_generator_spatio.py) -> LTN epoch (MIL + CE on those labels,
    Train/temporal_transformer_shanghaitech.py) -> LTN pseudo labels, thr 0.65
    (Train/pseudo_labels_generator_temporal.py).  Both TrainSteps must own their optimizer.  Returns the two label
    dicts, the last loss of each epoch and the wall-clock seconds of the four phases (device synchronised)."""
    import time
    import numpy as np
    rng = rng if rng is not None else np.random.RandomState(0)
    swl, lwl = stn.wl, ltn.wl
    steps = steps_per_epoch or max(1, len(corpus) // swl.batch_size)
    abn = dict(zip(corpus.abn_keys, corpus.abn))
    t = {}

    def clock(name, t0):
        torch.cuda.synchronize()
        t[name] = time.perf_counter() - t0

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stn.encoder.train(); stn.head.train()
    saved = corpus.pseudo
    corpus.pseudo = None
    for _ in range(steps):
        feats, _ = corpus.sample_batch(swl.batch_size, swl.part_num, swl.part_len, "uniform", rng)
        stn.zero_grad()
        s_terms = stn.forward_backward(feats.view(-1, swl.n_patch, swl.d_model), None, swl.batch_size)
    clock("stn_epoch", t0)
    t0 = time.perf_counter()
    stn_labels = score_videos(stn.encoder, stn.head, abn, part_len=1, threshold=thr_stn, n_patch=swl.n_patch,
                              cls_fast_path=cls_fast_path)
    clock("stn_labels", t0)
    t0 = time.perf_counter()
    corpus.pseudo = stn_labels
    ltn.encoder.train(); ltn.head.train()
    for _ in range(steps):
        feats, labs = corpus.sample_batch(lwl.batch_size, lwl.part_num, lwl.part_len, "uniform", rng)
        ltn.zero_grad()
        l_terms = ltn.forward_backward(feats, labs, lwl.batch_size)
    clock("ltn_epoch", t0)
    t0 = time.perf_counter()
    ltn_labels = score_videos(ltn.encoder, ltn.head, abn, part_len=lwl.part_len, threshold=thr_ltn, n_patch=lwl.n_patch,
                              cls_fast_path=cls_fast_path)
    clock("ltn_labels", t0)
    corpus.pseudo = saved
    n_clips = sum(int(v.shape[0]) for v in corpus.abn)
    return dict(stn_labels=stn_labels, ltn_labels=ltn_labels, stn_loss=s_terms["loss"].item(), ltn_loss=l_terms["loss"].item(),
                seconds=t, steps_per_epoch=steps, abnormal_clips=n_clips,
                stn_label_windows=n_clips, ltn_label_windows=sum(-(-int(v.shape[0]) // lwl.part_len) for v in corpus.abn))
