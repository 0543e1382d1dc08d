"""Sharded pseudo-label generation on real ranks: two processes (gloo rendezvous, both on cuda:0 - the data path has no
collective, so one GPU is enough) each score their shard of a synthetic corpus; the dict rank 0 merges must equal the
single-process sweep BIT FOR BIT, and the label file must have the reference's format
(Train/pseudo_labels_generator_temporal.py:113-145)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

KW = dict(n_layers=2, n_head=2, d_k=64, d_v=64, d_model=128, d_inner=256, MHA_layerNorm=True, FFN_layerNorm=True,
          weight_init=False, relative_pe=True, window_size=4, window_depth=3)
N_VIDEOS, PART_LEN, N_PATCH, THR = 26, 3, 16, 0.45


def _models(dev):
    from lstc_vad_b200.models import Classifier, Encoder
    torch.manual_seed(5)
    return Encoder(**KW).to(dev).eval(), Classifier(128, 0.6).to(dev).eval()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lstc_vad_b200.harness import merge_label_dicts, save_pseudo_labels, sharded_label_sweep
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    enc, cls = _models(dev)
    scores, st = sharded_label_sweep(enc, cls, N_VIDEOS, PART_LEN, N_PATCH, world, rank, threshold=THR, seed=9,
                                     max_windows=64)
    assert st["videos"] in (N_VIDEOS // world, N_VIDEOS // world + 1)
    merged = merge_label_dicts(scores, dist.group.WORLD)
    if rank == 0:
        save_pseudo_labels(out, merged)
        torch.save(merged, out + ".pt")
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_label_sweep_equals_single_rank_bit_for_bit(tmp_path):
    from lstc_vad_b200.harness import merge_label_dicts, sharded_label_sweep, synthetic_video_lengths
    out = str(tmp_path / "labels.npy")
    port = 29900 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out + ".pt", weights_only=False)
    dev = torch.device("cuda", 0)
    enc, cls = _models(dev)
    ref, st = sharded_label_sweep(enc, cls, N_VIDEOS, PART_LEN, N_PATCH, 1, 0, threshold=THR, seed=9, max_windows=4096)
    ref = merge_label_dicts(ref)
    lengths = synthetic_video_lengths(N_VIDEOS, 9)
    assert list(got) == list(ref) == [f"v{i:05d}" for i in range(N_VIDEOS)]
    n_kept = 0
    for i, k in enumerate(ref):
        assert got[k].shape == (lengths[i],)
        assert torch.equal(got[k], ref[k]), k          # bit for bit, although the forward batches differ
        n_kept += int((ref[k] > 0).sum())
    assert 0 < n_kept < sum(lengths)                    # the threshold keeps some clips and zeroes others
    # the reference's on-disk format: np.save of {key + '.npy': float32 [n_clips, 1]}
    disk = np.load(out, allow_pickle=True).item()
    assert set(disk) == {k + ".npy" for k in ref}
    for i, k in enumerate(ref):
        a = disk[k + ".npy"]
        assert a.dtype == np.float32 and a.shape == (lengths[i], 1)
        assert np.array_equal(a[:, 0], ref[k].numpy())
