"""Diagnostic: at a BASELINE config's full size, compare per layer  (a) K1-bwd class sums added over classes,
(b) autograd's bn.weight.grad (cuDNN's fp32 reduction) and (c) an fp64 torch evaluation of sum dy * xhat on the very
tensors the hooks saw.  Prints the error of (a) and of (b) against (c) relative to the absolute mass sum |dy * xhat|.

    python scripts/check_full_size.py --config c2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dcfp_b200.scorer import CalibrationRun
from dcfp_b200.workloads.segnets import CONFIGS, build_segnet
from dcfp_b200.workloads.synthetic import synthetic_batch


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--config", default="c2")
    p.add_argument("--every", type=int, default=8, help="check every n-th scored layer in fp64 (all of them costs memory)")
    a = p.parse_args()
    c = CONFIGS[a.config]
    K = c["num_classes"]
    model = build_segnet(c["arch"], c["backbone"], K, seed=0).cuda().to(memory_format=torch.channels_last)
    x, y = synthetic_batch([0, 1], K, c["height"], c["width"])
    x = x.cuda().contiguous(memory_format=torch.channels_last)
    y = y.cuda()
    run = CalibrationRun(model, K, r=0.999, seed=0, keep_totals=True)
    sc = run.scorer
    watched = {n: m for i, (n, m) in enumerate(sc.layers) if i % a.every == 0 or i < 3}
    seen = {}
    handles = []
    for n, m in watched.items():
        def hook(mod, inp, out, n=n):
            xs = inp[0].detach()
            seen[n] = {"x": xs}
            out.register_hook(lambda g, n=n: seen[n].__setitem__("dy", g.detach()))
        handles.append(m.register_forward_hook(hook))
    run.step(x, y, mb_index=0)
    torch.cuda.synchronize()
    S1 = sc.totals[0].sum(0)
    print("%-34s %5s %9s %11s %11s %11s %11s" % ("layer", "C", "h x w", "mean|g|", "mass", "K1 err/mass", "cuDNN err/mass"))
    worst = 0.0
    for n, m, lo, hi in zip(sc.names, [m for _, m in sc.layers], sc.offsets[:-1], sc.offsets[1:]):
        if n not in seen:
            continue
        xs, dy = seen[n]["x"].double(), seen[n]["dy"].double()
        mean = xs.mean(dim=(0, 2, 3), keepdim=True)
        var = xs.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        v = dy * (xs - mean) * torch.rsqrt(var + m.eps)
        ref = v.sum(dim=(0, 2, 3))
        mass = v.abs().sum(dim=(0, 2, 3)) + 1e-300
        k1 = ((S1[lo:hi] - ref).abs() / mass).max().item()
        cd = ((m.weight.grad.double() - ref).abs() / mass).max().item()
        worst = max(worst, k1)
        print("%-34s %5d %4dx%-4d %11.4g %11.4g %11.3g %11.3g" % (n, hi - lo, xs.shape[2], xs.shape[3], ref.abs().mean().item(),
                                                                 mass.mean().item(), k1, cd))
    for h in handles:
        h.remove()
    run.close()
    print("worst K1 error / mass: %.3g" % worst)


if __name__ == "__main__":
    main()
