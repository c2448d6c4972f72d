"""Calibration scoring -> score.pth -> FLOPs-ratio search -> pruned.pth + channel_cfg.pth, all on the GPU path.

    python examples/score_and_prune.py --config c1 --images 16 --prune-ratio 0.5 --out /tmp/dcfp_out

Uses the synthetic calibration set and the random-init workload nets (there is no dataset on the GPU box); with a real
model / data loader, pass your own `model`, `images`, `labels` to `score_calibration_set`.
The three files it writes are the reference's formats (train.py:286-287, prune.py:97-98)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dcfp_b200.pruners import formats
from dcfp_b200.pruners.search import prune_to_flops_ratio
from dcfp_b200.scorer import score_calibration_set
from dcfp_b200.workloads.segnets import CONFIGS, build_segnet
from dcfp_b200.workloads.synthetic import synthetic_batch


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    p.add_argument("--images", type=int, default=16)
    p.add_argument("--prune-ratio", type=float, default=0.5)
    p.add_argument("--out", default="./dcfp_out")
    a = p.parse_args()
    c = CONFIGS[a.config]
    os.makedirs(a.out, exist_ok=True)
    model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0).cuda()
    images, labels = synthetic_batch(list(range(a.images)), c["num_classes"], c["height"], c["width"])
    t = time.time()
    out = score_calibration_set(model, images, labels, c["num_classes"], micro_batch=2, r=0.999)
    torch.cuda.synchronize()
    print("scored %d images in %.2f s (%d kernels of this library)" % (a.images, time.time() - t, out["_stats"]["launches"]))
    score = os.path.join(a.out, "score.pth")
    formats.save_score(out["eic"], score)
    model.criterion = None
    t = time.time()
    sub, channel_cfg, gp = prune_to_flops_ratio(model.cpu(), score, prune_ratio=a.prune_ratio)
    print("pruned at global_percent %.2f in %.2f s" % (gp, time.time() - t))
    formats.save_pruned(sub, channel_cfg, a.out)
    # what a consumer does (prune.py:100-110, train.py:200-207): fresh model -> channel_cfg sizes -> pruned weights
    fresh = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=1)
    formats.load_pruned_model(fresh, os.path.join(a.out, "channel_cfg.pth"), os.path.join(a.out, "pruned.pth"))
    kept = sum(v["out_channels"] for v in channel_cfg.values())
    raw = sum(v["raw_out_channels"] for v in channel_cfg.values())
    print("kept %d of %d output channels; wrote %s" % (kept, raw, a.out))


if __name__ == "__main__":
    main()
