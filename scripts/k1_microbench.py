"""Quick K1 timing on resident feature maps (development aid; bench.py is the judged entry)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcfp_b200 import ops
from dcfp_b200.workloads.synthetic import synthetic_batch

ops.require_gpu()
dev = torch.device("cuda")
K = int(os.environ.get("K", 19))
mb = 2
H0, W0 = (512, 1024) if K == 19 else (512, 512)
_, lab = synthetic_batch([0, 1], K, H0, W0, fragmentation=os.environ.get("FRAG") or None)
lab = lab.to(dev)
KEYS = {}
def keys_for(shapes):
    out = []
    for c, h, w in shapes:
        if (h, w) not in KEYS:
            KEYS[(h, w)] = ops.label_keys(lab, h, w, K)
        out.append(KEYS[(h, w)])
    return out
s = 1 if K == 19 else 2
shapes = [(1024, 64, 128 // s)] * 24 + [(256, 64, 128 // s)] * 52 + [(512, 64, 128 // s)] * 12 + [(2048, 64, 128 // s)] * 3 + \
    [(256, 128, 256 // s)] * 4 + [(64, 256, 512 // s)] * 2 + [(128, 256, 512 // s)] + [(64, 128, 256 // s)] * 6 + \
    [(128, 64, 128 // s)] * 7 + [(128, 128, 256 // s)] + [(256, 1, 1)]


NHWC = os.environ.get("LAYOUT", "nchw") == "nhwc"
fmt = torch.channels_last if NHWC else torch.contiguous_format


def run(dtype, bwd, iters=5):
    xs = [torch.randn(mb, c, h, w, device=dev).to(dtype).contiguous(memory_format=fmt) for c, h, w in shapes]
    dys = [torch.randn_like(x) for x in xs] if bwd else None
    sc = [torch.ones(c, device=dev) for c, _, _ in shapes] if bwd else None
    sf = [torch.zeros(c, device=dev) for c, _, _ in shapes] if bwd else None
    S1 = [torch.zeros(K, c, dtype=torch.float64, device=dev) for c, _, _ in shapes]
    S2 = [torch.zeros_like(t) for t in S1]
    kl = keys_for(shapes)
    nbytes = sum(x.numel() * x.element_size() for x in xs) * (2 if bwd else 1)
    for _ in range(3):
        ops.class_stats_grouped(xs, kl, K, S1, S2, dys=dys, scales=sc, shifts=sf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.class_stats_grouped(xs, kl, K, S1, S2, dys=dys, scales=sc, shifts=sf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(("nhwc " if NHWC else "nchw ") + "K=%d %s %s grouped: %.3f ms  %.1f GB/s (%.1f%% of 6551)  %.0f img/s" %
          (K, str(dtype).split(".")[-1], "bwd" if bwd else "fwd", ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.514, mb / ms * 1e3), flush=True)
    # per-layer launches, the dominant shape
    x = xs[30]
    a, b = S1[30], S2[30]
    for _ in range(3):
        ops.class_stats(x, kl[30], K, a, b, dy=None if not bwd else dys[30], scale=None if not bwd else sc[30], shift=None if not bwd else sf[30])
    e0.record()
    for _ in range(50):
        ops.class_stats(x, kl[30], K, a, b, dy=None if not bwd else dys[30], scale=None if not bwd else sc[30], shift=None if not bwd else sf[30])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    nb = x.numel() * x.element_size() * (2 if bwd else 1)
    print("   single layer %s: %.1f us  %.1f GB/s" % (tuple(x.shape), ms * 1e3, nb / ms / 1e6), flush=True)
    del xs, dys


for dtype in (torch.float32, torch.bfloat16):
    for bwd in (False, True):
        run(dtype, bwd)
# reference point: plain device copy
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
print("copy 2x2GiB: %.1f GB/s" % (2 * a.numel() * 2 / (e0.elapsed_time(e1) / 5) / 1e6))
