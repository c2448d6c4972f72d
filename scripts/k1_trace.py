"""Per-phase timeline of the fused K1 kernel (B1): needs a trace build -- DCFP_K1_TRACE=1 python -m dcfp_b200.build -- whose CTAs
stamp %globaltimer at seven points (csrc/k1_nhwc.cuh: K1_TRACE).  Prints mean / max over the CTAs of one launch per shape."""
import ctypes, sys
sys.path.insert(0, ".")
import numpy as np, torch
from dcfp_b200 import ops, abi
from dcfp_b200.workloads.synthetic import synthetic_batch
lib = ctypes.CDLL(abi.LIB_PATH)
lib.dcfp_debug_k1_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
_, lab = synthetic_batch([0, 1], 19, 512, 1024, fragmentation="street")
lab = lab.to(dev)
for C, h, w in [(256, 64, 128), (1024, 64, 128), (64, 128, 256)]:
    xs = [torch.randn(2, C, h, w, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(6)]
    dys = [torch.randn_like(x) * 1e-3 for x in xs]
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    keys = ops.label_keys(lab, h, w, 19)
    R = 20
    S1 = torch.zeros(R, C, dtype=torch.float32, device=dev); S2 = torch.zeros_like(S1)
    sums = ops.bn_scratch(C, dev)
    y, mean, invstd = ops.bn_forward(xs[0], gamma, beta, None, None, sums, 0.1, 1e-5, True)
    res = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ops.bn_backward(xs[i % 6], dys[i % 6], gamma, beta, mean, invstd, keys, S1, S2, R, sums, True, True, phases=1)
        e1.record()
        torch.cuda.synchronize()
        buf = np.zeros(160 * 8, dtype=np.uint64)
        lib.dcfp_debug_k1_trace(buf.ctypes.data_as(ctypes.c_void_p))
        t = buf.reshape(160, 8)[:148, :7].astype(np.int64)
        t = t[t[:, 0] > 0]
        res.append((e0.elapsed_time(e1) * 1e3, t))
    ev, t = res[-1]
    t0 = t[:, 0].min()
    names = ["start->setup", "setup->first data", "first data->loop end", "loop end->merged", "merged->totals", "totals->rows flushed"]
    d = np.diff(t, axis=1) / 1e3
    print("C=%d %dx%d: event %.1f us; CTAs %d; first CTA start spread %.1f us; kernel span (first start -> last end) %.1f us" % (
        C, h, w, ev, len(t), (t[:, 0].max() - t0) / 1e3, (t[:, 6].max() - t0) / 1e3))
    for j, n in enumerate(names):
        print("   %-26s mean %6.2f us  max %6.2f us" % (n, d[:, j].mean(), d[:, j].max()))
