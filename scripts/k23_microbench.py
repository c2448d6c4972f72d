"""K2 (thresholds + masks) and K3 (grouped channel gather, bias compensation) timings on real model tensors, and the
whole DCFPPruner.prune_model call (development aid; numbers are copied into profiles/ and DESIGN.md)."""
import copy
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dcfp_b200 import ops
from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
from dcfp_b200.workloads.segnets import CONFIGS, build_segnet

ops.require_gpu()
dev = torch.device("cuda")
cfg = os.environ.get("CFG", "c2")
c = CONFIGS[cfg]
model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0, with_loss=False)
rng = np.random.RandomState(1)
eic = {n: torch.from_numpy(rng.rand(m.weight.numel()).astype(np.float32)) for n, m in model.named_modules()
       if isinstance(m, torch.nn.BatchNorm2d) and n not in model.ignore_prune_layer}
score = tempfile.NamedTemporaryFile(suffix=".pth", delete=False).name
torch.save({"eic": eic}, score)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---- K2 on the concatenated score vector of this model
names = list(eic.keys())
sizes = [eic[n].numel() for n in names]
offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
groups = torch.tensor([0 if n.startswith("backbone") else 1 for n in names], dtype=torch.int32, device=dev)
mk = torch.tensor([max(int(s * 0.02), 1) for s in sizes], dtype=torch.int32, device=dev)
sv = torch.cat([eic[n] for n in names]).to(dev)
k = [int(sum(s for s, n in zip(sizes, names) if n.startswith("backbone") == (g == 0)) * 0.5) for g in (0, 1)]
ms = timed(lambda: ops.thresh_mask(sv, offs, groups, mk, k[0], k[1]))
print("%s K2 thresh+mask over %d channels / %d layers: %.1f us per call (2 launches)" % (cfg, sv.numel(), len(names), ms * 1e3))
g = torch.randn_like(sv)
w = torch.ones_like(sv)
e = torch.zeros_like(sv)
ms = timed(lambda: ops.eic_update_flat(g, w, e, 0.999, False))
print("%s K2 EIC update (flat, %d channels): %.1f us" % (cfg, sv.numel(), ms * 1e3))
K = c["num_classes"]
step = torch.zeros(2, K, sv.numel(), dtype=torch.float64, device=dev)
tot = torch.zeros_like(step)
ms = timed(lambda: ops.fold_step(step, tot))
print("%s fold_step over a [2,%d,%d] fp64 arena (%.0f MB): %.1f us" % (cfg, K, sv.numel(), step.numel() * 8 / 1e6, ms * 1e3))

# ---- K3: gather every conv weight of the model with ~50 % of out/in channels kept (one grouped launch)
ws = [m.weight.detach().to(dev).contiguous() for m in model.modules() if isinstance(m, torch.nn.Conv2d)]
oi, ii = [], []
for t in ws:
    o = torch.nonzero(torch.rand(t.shape[0], device=dev) > 0.5).reshape(-1).to(torch.int32)
    i = torch.nonzero(torch.rand(t.shape[1], device=dev) > 0.5).reshape(-1).to(torch.int32) if t.shape[1] > 3 else None
    oi.append(o)
    ii.append(i)
outs = ops.channel_gather_grouped(ws, oi, ii)
rd = sum(t.numel() * 4 for t in ws)
wr = sum(t.numel() * 4 for t in outs)
ms = timed(lambda: ops.channel_gather_grouped(ws, oi, ii), iters=10)
print("%s K3 grouped gather of %d conv weights: %.1f MB source, %.1f MB kept: %.3f ms (incl. output allocation + table upload)"
      % (cfg, len(ws), rd / 1e6, wr / 1e6, ms))
print("   useful bytes (kept read + kept written) %.1f MB -> %.1f GB/s; source-sweep bytes %.1f MB -> %.1f GB/s" %
      (2 * wr / 1e6, 2 * wr / ms / 1e6, (rd + wr) / 1e6, (rd + wr) / ms / 1e6))
big = max(ws, key=lambda t: t.numel())
act = torch.rand(big.shape[1], device=dev)
ms = timed(lambda: ops.bias_comp(big, act))
print("%s bias_comp on %s: %.1f us = %.1f GB/s" % (cfg, tuple(big.shape), ms * 1e3, big.numel() * 4 / ms / 1e6))

# ---- the whole prune_model call (reference on the CPU: 15-21 s)
for where in ("cpu", "cuda"):
    for rep in range(2):
        m = copy.deepcopy(model).to(where)
        torch.cuda.synchronize()
        t = time.time()
        pr = DCFPPruner(global_percent=0.5, layer_keep=0.02, score_file=score)
        sub, ccfg = pr.prune_model(m, except_start_keys=["conv_deepsup"])
        torch.cuda.synchronize()
        print("%s prune_model, model on %s, call %d: %.2f s (%d launches so far)" % (cfg, where, rep, time.time() - t, ops.launch_count()))
