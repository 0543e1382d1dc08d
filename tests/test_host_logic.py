"""CPU-only tests of the host-side logic: state_dict layout / seeded init of the module mirror against the
reference-generated goldens, workload FLOP accounting, window policies, video sharding, CPU-tensor rejection."""
from pathlib import Path

import pytest
import torch

from oracle import lstc_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name,head", [("ltn_relpe", "cls"), ("ltn_ucf_sliced", "cls"), ("stn_plain", "reg"),
                                       ("stn_relpe2d_cls_pos", "reg")])
def test_state_dict_layout_matches_reference_checkpoints(name, head):
    from lstc_vad_b200.models import Classifier, Encoder, Regressor
    c = torch.load(GOLD / f"{name}.pt", weights_only=False)
    enc = Encoder(**c["enc_kwargs"])
    sd = enc.state_dict()
    ref = c["enc_state"]
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        if k.endswith("relative_position_index"):
            assert torch.equal(sd[k], ref[k])  # integer buffer: bit-exact
    enc.load_state_dict(ref, strict=True)
    D = c["enc_kwargs"]["d_model"]
    h = Classifier(D) if head == "cls" else Regressor(D)
    href = c["cls_state"] if head == "cls" else c["reg_state"]
    assert list(h.state_dict().keys()) == list(href.keys())
    h.load_state_dict(href, strict=True)
    # DataParallel-style prefixed checkpoints load after the consumers' k[7:] strip
    pref = {"module." + k: v for k, v in ref.items()}
    enc.load_state_dict({k[7:]: v for k, v in pref.items()}, strict=True)


def test_seeded_init_reproduces_the_golden_weights():
    """Same seed -> same parameters as the reference constructor (RNG consumed in the same order)."""
    from lstc_vad_b200.models import Classifier, Encoder
    c = torch.load(GOLD / "ltn_relpe.pt", weights_only=False)
    torch.manual_seed(0)  # seed used by oracle/make_golden.py for this case
    enc = Encoder(**c["enc_kwargs"])
    cls = Classifier(c["enc_kwargs"]["d_model"], 0.6, weight_init=True)
    for k, v in enc.state_dict().items():
        assert torch.equal(v, c["enc_state"][k]), k
    for k, v in cls.state_dict().items():
        assert torch.equal(v, c["cls_state"][k]), k


def test_parameters_without_gradient_exist_in_state_dict():
    from lstc_vad_b200.models import Encoder
    enc = Encoder(n_layers=1, n_head=1, d_k=64, d_v=64, d_model=64, d_inner=64, MHA_layerNorm=False, FFN_layerNorm=False)
    keys = set(enc.state_dict())
    assert {"layer_norm.weight", "layer_stack.0.slf_attn.layer_norm.bias", "layer_stack.0.pos_ffn.layer_norm.weight"} <= keys


def test_modules_refuse_cpu_tensors():
    from lstc_vad_b200.models import Classifier, Encoder
    enc = Encoder(n_layers=1, n_head=1, d_k=64, d_v=64, d_model=64, d_inner=64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(2, 4, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Classifier(64)(torch.zeros(2, 64))


def test_workload_flop_accounting_matches_survey():
    from lstc_vad_b200.harness import WORKLOADS
    assert abs(WORKLOADS["ltn_sht"].fwd_flops_per_window() / 1e9 - 9.926) < 0.01
    assert abs(WORKLOADS["ltn_ucf"].fwd_flops_per_window() / 1e9 - 3.836) < 0.01
    assert abs(WORKLOADS["ltn_ubnormal"].fwd_flops_per_window() / 1e9 - 8.316) < 0.01
    assert abs(WORKLOADS["stn_sht"].fwd_flops_per_window() / 1e9 - 2.985) < 0.01
    assert WORKLOADS["ltn_sht"].windows_per_step == 1280 and WORKLOADS["ltn_ucf"].windows_per_step == 2560
    assert WORKLOADS["stn_sht"].windows_per_step == 8960


def test_window_policies_match_oracle_and_reference_loops():
    from lstc_vad_b200.harness import video_windows
    for n in range(1, 40):
        for T in (1, 2, 3, 5):
            for bs in (False, True):
                mine = [(b, e) for b, e, _ in video_windows(n, T, bs)]
                assert mine == O.window_bounds(n, T, bs), (n, T, bs)
                assert sum(c for _, _, c in video_windows(n, T, bs)) == n  # every clip scored exactly once
    assert video_windows(7, 3, True) == [(0, 3, 3), (3, 6, 3), (4, 7, 1)]
    assert video_windows(7, 3, False) == [(0, 3, 3), (3, 6, 3), (6, 7, 1)]


def test_shard_videos_partitions_and_balances():
    from lstc_vad_b200.harness import shard_videos
    g = torch.Generator().manual_seed(0)
    n_clips = torch.randint(5, 200, (101,), generator=g).tolist()
    keys = [f"v{i:03d}" for i in range(101)]
    seen, loads = [], []
    for r in range(8):
        mine = shard_videos(keys, n_clips, 8, r)
        seen += mine
        loads.append(sum(n_clips[i] for i in mine))
    assert sorted(seen) == list(range(101))
    assert max(loads) - min(loads) <= max(n_clips)


def test_soft_clip_labels_match_oracle():
    from lstc_vad_b200.losses import soft_clip_labels
    g = torch.Generator().manual_seed(1)
    pseudo = torch.rand(5, 12, 1, generator=g)
    assert torch.equal(soft_clip_labels(pseudo, 5, 4, 3), O.soft_labels(pseudo, 5, 4, 3))


def test_dropout_stream_is_deterministic_and_distinct():
    from lstc_vad_b200 import functional as Fn
    Fn.set_dropout_stream(42, 0)
    a = [Fn.next_dropout(0.2, True) for _ in range(3)]
    Fn.set_dropout_stream(42, 0)
    b = [Fn.next_dropout(0.2, True) for _ in range(3)]
    assert a == b and len({x[2] for x in a}) == 3
    assert Fn.next_dropout(0.2, False) == Fn.NO_DROPOUT and Fn.next_dropout(0.0, True) == Fn.NO_DROPOUT


def test_weight_cache_is_not_fooled_by_recycled_parameters():
    """A new Parameter can reuse the id() / storage address / version of a dead one; the cache must miss."""
    from lstc_vad_b200.functional import WeightCache
    cache = WeightCache()
    built = []

    def get(p):
        return cache._get(("bf16", id(p), None, 0, 0), (p,), lambda: built.append(1) or float(p.sum()))

    p = torch.nn.Parameter(torch.ones(4, 4))
    assert get(p) == 16.0 and get(p) == 16.0 and len(built) == 1      # second call is a hit
    key_id, ptr = id(p), p.data_ptr()
    del p
    for _ in range(64):  # try to get an object with the same id / storage
        q = torch.nn.Parameter(torch.full((4, 4), 2.0))
        if id(q) == key_id or q.data_ptr() == ptr:
            break
    assert get(q) == 32.0                                             # never the stale 16.0
    with torch.no_grad():
        q.add_(1.0)                                                   # optimizer-style in-place update bumps _version
    assert get(q) == 48.0


def test_window_plan_is_video_windows_regrouped():
    """The batched scorer's plan (one view for all full windows + the ragged tail) covers exactly the windows of the
    reference loops (oracle.window_bounds), for both trailing-window policies."""
    from lstc_vad_b200.harness import video_windows, window_plan
    from oracle import lstc_oracle as O
    for T in (1, 2, 3, 5, 7):
        for n in range(0, 40):
            for backshift in (False, True):
                n_full, tail = window_plan(n, T, backshift)
                wins = [(i * T, (i + 1) * T) for i in range(n_full)] + ([tail[:2]] if tail else [])
                if n:
                    assert wins == O.window_bounds(n, T, backshift=backshift), (n, T, backshift)
                covers = [T] * n_full + ([tail[2]] if tail else [])
                assert covers == [c for _, _, c in video_windows(n, T, backshift)], (n, T, backshift)
                assert sum(covers) == n


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one) needs no GPU: one JSON line with
    the metric / unit of the native arm, `impl`, a `cpu_baseline` describing the run and a zero-copy `e2e`; under
    torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "windows/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("windows/sec fwd+bwd") and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    other = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                           capture_output=True, text=True, timeout=600, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=root)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_dropin_package_shadows_the_reference_module_paths():
    """With `dropin/` first on sys.path the reference scripts' own import lines (`from models.Encoder import Encoder`,
    Train/temporal_transformer_shanghaitech.py:16-18) resolve to the B200 mirror; `utils` is left to the reference."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "from models.Encoder import Encoder\n"
        "from models.EncoderLayer import EncoderLayer\n"
        "from models.MultiHeadAttention import MultiHeadAttention, ScaledDotProductAttention\n"
        "from models.FFN import PositionwiseFeedForward\n"
        "from models.PatchEmbedding import PatchEmbedding\n"
        "from models.Classifier import Classifier\n"
        "from models.Regressor import Regressor\n"
        "import importlib.util\n"
        "mods = {c.__module__ for c in (Encoder, EncoderLayer, MultiHeadAttention, PositionwiseFeedForward, PatchEmbedding, Classifier, Regressor)}\n"
        "assert all(m.startswith('lstc_vad_b200.models.') for m in mods), mods\n"
        "assert importlib.util.find_spec('utils') is None or 'dropin' not in (importlib.util.find_spec('utils').origin or '')\n"
        "e = Encoder(n_layers=1, n_head=2, d_k=64, d_v=64, d_model=64, d_inner=128, relative_pe=True, window_size=4, window_depth=3)\n"
        "assert 'layer_stack.0.slf_attn.relative_position_index' in e.state_dict()\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "dropin") + os.pathsep + root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_wgrad_split_fills_the_last_wave_of_the_persistent_grid():
    """Weight-gradient split-K plan (functional._wgrad_split): the persistent CTA-pair GEMM runs work items = output tiles x
    splits in waves of 74 tile pairs, and the smallest split whose last wave is >= 95 % full is taken, with at least four
    64-row k-blocks per split."""
    from lstc_vad_b200.functional import _wgrad_split
    rows = 1280 * 49
    expected = {(2048, 2048): 8, (4096, 2048): 4, (2048, 4096): 4, (6144, 2048): 3}   # the four weight shapes of an LTN layer
    for (n_out, k_out), s in expected.items():
        assert _wgrad_split(n_out, k_out, rows) == s
        tiles = ((n_out + 255) // 256) * ((k_out + 255) // 256)
        items = tiles * s
        assert items / (-(-items // 74) * 74) >= 0.95
    # short reductions are never split below four k-blocks per split, tiny outputs use the single-CTA kernel's 148 units
    assert _wgrad_split(2048, 2048, 256) == 1
    assert 1 <= _wgrad_split(32, 512, 1280) <= (1280 // 64) // 4


def test_bench_reads_the_committed_gemm_traffic_figure():
    """`roofline.traffic` is the mean DRAM bytes per GEMM launch of the committed ncu launch list (not a constant in
    bench.py), and only claimed for the workload that list was captured on."""
    import bench
    t = bench.committed_gemm_traffic("ltn_sht")
    tail = (Path(__file__).resolve().parent.parent / "profiles" / "r2_step_launches_traffic.txt").read_text().strip().splitlines()[-1]
    assert t is not None and f"{t / 1e9:.3f} GB" in tail
    assert 0.8e9 <= t <= 1.3e9            # 0.80 GB algorithmic + the weight-gradient re-reads
    assert bench.committed_gemm_traffic("stn_sht") is None
