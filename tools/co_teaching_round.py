"""One synthetic co-teaching round on this GPU: STN epoch -> STN labels -> LTN epoch -> LTN
labels, on a corpus generated in HBM.  Prints one JSON line with the seconds per phase.

    python tools/co_teaching_round.py [--videos 400] [--cls-fast-path]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstc_vad_b200.harness import WORKLOADS, DeviceCorpus, TrainStep, co_teaching_round, synthetic_corpus  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=400)
    ap.add_argument("--cls-fast-path", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    swl, lwl = WORKLOADS["stn_sht"], WORKLOADS["ltn_sht"]
    normal, abnormal = synthetic_corpus(args.videos, lwl.n_patch, lwl.d_model, dev)
    corpus = DeviceCorpus(normal, abnormal, dev)
    stn = TrainStep(swl, dev, seed=0, optimizer=True, cls_fast_path=args.cls_fast_path)
    ltn = TrainStep(lwl, dev, seed=1, optimizer=True, cls_fast_path=args.cls_fast_path)
    co_teaching_round(corpus, stn, ltn, steps_per_epoch=1, rng=np.random.RandomState(1))   # warm-up (caches, allocator)
    out = co_teaching_round(corpus, stn, ltn, rng=np.random.RandomState(2), cls_fast_path=args.cls_fast_path)
    sec = out["seconds"]
    line = {"videos": args.videos, "abnormal_clips": out["abnormal_clips"], "steps_per_epoch": out["steps_per_epoch"],
            "seconds": {k: round(v, 4) for k, v in sec.items()}, "round_seconds": round(sum(sec.values()), 4),
            "stn_label_windows_per_s": round(out["stn_label_windows"] / sec["stn_labels"]),
            "ltn_label_windows_per_s": round(out["ltn_label_windows"] / sec["ltn_labels"]),
            "stn_train_windows_per_s": round(swl.windows_per_step * out["steps_per_epoch"] / sec["stn_epoch"]),
            "ltn_train_windows_per_s": round(lwl.windows_per_step * out["steps_per_epoch"] / sec["ltn_epoch"]),
            "cls_fast_path": args.cls_fast_path, "stn_loss": out["stn_loss"], "ltn_loss": out["ltn_loss"]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
