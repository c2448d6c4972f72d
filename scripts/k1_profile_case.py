"""Small fixed K1 case for ncu (fp32 fwd grouped, ~0.8 GB of resident maps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcfp_b200 import ops
from dcfp_b200.workloads.synthetic import synthetic_batch

ops.require_gpu()
dev = torch.device("cuda")
K = int(os.environ.get("K", 19))
bwd = os.environ.get("BWD", "0") == "1"
dtype = torch.bfloat16 if os.environ.get("DT", "f32") == "bf16" else torch.float32
S = 1 if K == 19 else 2  # 512x1024 (Cityscapes) vs 512x512 (ADE / COCO-Stuff) label maps
_, lab = synthetic_batch([0, 1], K, 512, 1024 // S, fragmentation=os.environ.get("FRAG") or None)
lab = lab.to(dev)
KEYS = {}
def keys_for(shapes):
    out = []
    for c, h, w in shapes:
        if (h, w) not in KEYS:
            KEYS[(h, w)] = ops.label_keys(lab, h, w, K)
        out.append(KEYS[(h, w)])
    return out
shapes = [(1024, 64, 128 // S)] * 5 + [(256, 64, 128 // S)] * 12 + [(512, 64, 128 // S)] * 3 + [(2048, 64, 128 // S)] + \
    [(256, 128, 256 // S)] + [(64, 256, 512 // S)]
fmt = torch.channels_last if os.environ.get("LAYOUT", "nchw") == "nhwc" else torch.contiguous_format
xs = [torch.randn(2, c, h, w, device=dev).to(dtype).contiguous(memory_format=fmt) for c, h, w in shapes]
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
for _ in range(3):
    ops.class_stats_grouped(xs, kl, K, S1, S2, dys=dys, scales=sc, shifts=sf)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("bytes %.1f MB  %.3f ms  %.1f GB/s" % (nbytes / 1e6, ms, nbytes / ms / 1e6))
