"""Per-layer timing of the fused BN passes on the c2 layer shapes (development aid): F1 (K1, one class), F2 (apply), B1 (fused
class-keyed reduction), B2 (dx).  Buffers rotate through > 126 MB so that the first pass of a pair reads HBM, as in a step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dcfp_b200 import ops  # noqa: E402
from dcfp_b200.workloads.synthetic import synthetic_batch  # noqa: E402

ops.require_gpu()
dev = torch.device("cuda")
K = int(os.environ.get("K", 19))
H0, W0 = 512, 1024
_, lab = synthetic_batch([0, 1], K, H0, W0)
lab = lab.to(dev)
SHAPES = [(256, 64, 128), (512, 64, 128), (1024, 64, 128), (2048, 64, 128), (64, 256, 512), (128, 256, 512), (64, 128, 256), (128, 64, 128)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
PEAK = 6551.4


GRAPH = os.environ.get("GRAPH", "1") == "1"  # replay the n calls from a CUDA graph: GPU time, not the host's launch rate


def timeit(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


for C, h, w in SHAPES:
    nbytes = 2 * C * h * w * 4
    nbuf = max(2, int(400e6 // (3 * nbytes)) + 1)  # rotate through > 3x L2 worth of (x, dy, out)
    xs = [torch.randn(2, C, h, w, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(nbuf)]
    dys = [torch.randn_like(x) * 1e-3 for x in xs]
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    keys = ops.label_keys(lab, h, w, K)
    R = K + 1
    S1 = torch.zeros(R, C, dtype=torch.float64, device=dev)
    S2 = torch.zeros_like(S1)
    sums = ops.bn_scratch(C, dev)
    y, mean, invstd = ops.bn_forward(xs[0], gamma, beta, None, None, sums, 0.1, 1e-5, True)
    one = torch.zeros(1, C, dtype=torch.float64, device=dev)
    one2 = torch.zeros_like(one)
    n = 40
    t_f1 = timeit(lambda i: ops.bn_forward(xs[i % nbuf], gamma, beta, None, None, sums, 0.1, 1e-5, True, phases=1), n)
    t_f = timeit(lambda i: ops.bn_forward(xs[i % nbuf], gamma, beta, None, None, sums, 0.1, 1e-5, True), n)
    # bottleneck tail: bn -> (+ shortcut) -> ReLU in the normalise pass, against the same three steps as three kernels
    t_fr = timeit(lambda i: ops.bn_forward(xs[i % nbuf], gamma, beta, None, None, sums, 0.1, 1e-5, True, residual=dys[i % nbuf]), n)

    def apart(i):
        z, _, _ = ops.bn_forward(xs[i % nbuf], gamma, beta, None, None, sums, 0.1, 1e-5, False)
        torch.relu_(z + dys[i % nbuf])
    t_fa = timeit(apart, n)
    wsp = ops.bn_workspace(C, dev)

    def coop(i):
        sums.zero_()
        ops.bn_forward(xs[i % nbuf], gamma, beta, None, None, sums, 0.1, 1e-5, True, workspace=wsp)
    t_z = timeit(lambda i: sums.zero_(), n)
    t_c = timeit(coop, n) - t_z
    t_b1 = timeit(lambda i: ops.bn_backward(xs[i % nbuf], dys[i % nbuf], gamma, beta, mean, invstd, keys, S1, S2, R, sums, True, True, phases=1), n)
    t_b = timeit(lambda i: ops.bn_backward(xs[i % nbuf], dys[i % nbuf], gamma, beta, mean, invstd, keys, S1, S2, R, sums, True, True), n)
    t_k1 = timeit(lambda i: ops.class_stats(xs[i % nbuf], keys, R, S1, S2, dy=dys[i % nbuf], scale=invstd, shift=mean, affine_mode=1), n)
    gb = lambda us, mult: mult * nbytes / us / 1e3
    print("C=%4d %3dx%3d %6.1f MB | F1 %6.1f us %5.0f GB/s | F1+F2 %6.1f us (alg 2x: %5.0f GB/s) | coopF %6.1f us (alg 2x: %5.0f GB/s) | B1 %6.1f us %5.0f GB/s | K1bwd %6.1f us | "
          "B1+B2 %6.1f us (alg 3x: %5.0f GB/s) | tail fused %6.1f us (alg 3x: %5.0f GB/s) vs bn, add, relu apart %6.1f us | ideal@peak F %.1f B %.1f us" %
          (C, h, w, nbytes / 1e6, t_f1, gb(t_f1, 1), t_f, gb(t_f, 2), t_c, gb(t_c, 2), t_b1, gb(t_b1, 2), t_k1, t_b, gb(t_b, 3), t_fr, gb(t_fr, 3), t_fa,
           2 * nbytes / PEAK / 1e3, 3 * nbytes / PEAK / 1e3), flush=True)
