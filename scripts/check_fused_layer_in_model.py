"""Inside a real scoring step: for a few fused BN layers capture (x, dy, dx) through tensor hooks and compare dx / dgamma /
dbeta with the fp64 factored oracle (oracle/bn_ref.py) on the same device tensors."""
import sys

import torch

sys.path.insert(0, ".")
from dcfp_b200.scorer import CalibrationRun  # noqa: E402
from dcfp_b200.workloads.segnets import build_segnet  # noqa: E402
from dcfp_b200.workloads.synthetic import synthetic_batch  # noqa: E402
from oracle import bn_ref  # noqa: E402

DEV = "cuda"
K, H, W = 19, 128, 256
torch.backends.cudnn.allow_tf32 = False
model = build_segnet("deeplabv3", "resnet50", K, seed=0).to(DEV).to(memory_format=torch.channels_last)
x, y = synthetic_batch([0, 1], K, H, W)
x, y = x.to(DEV).contiguous(memory_format=torch.channels_last), y.to(DEV)
run = CalibrationRun(model, K, seed=5, fused=True)
run.step(x, y, mb_index=0)
names = ["last_conv.4", "last_conv.1", "aspp.aspp1.bn", "backbone.layer4.2.bn3", "backbone.layer1.0.bn1", "backbone.bn1", "backbone.conv1.1"]
cap = {}
handles = []
mods = dict(model.named_modules())
for n in names:
    m = mods[n]

    def pre(mod, inp, n=n):
        xin = inp[0]
        cap[n] = {"x": xin.detach()}
        if xin.requires_grad:
            xin.register_hook(lambda g, n=n: cap[n].__setitem__("dx", g.detach().clone()))

    def post(mod, inp, out, n=n):
        cap[n]["tag"] = getattr(out, "_dcfp_bn", None)
        out.register_hook(lambda g, n=n: cap[n].__setitem__("dy", g.detach().clone()))

    handles.append(m.register_forward_pre_hook(pre))
    handles.append(m.register_forward_hook(post))
run.step(x, y, mb_index=0)
for n in names:
    c, m = cap[n], mods[n]
    xx, dy = c["x"].double(), c["dy"].double()
    relu = bool(c["tag"] and c["tag"][1])
    g, b = m.weight.detach().double(), m.bias.detach().double()
    _, mean, invstd = bn_ref.bn_relu_forward(xx, g, b, m.eps, relu)
    dx, dg, db = bn_ref.bn_relu_backward(xx, dy, g, b, mean, invstd, relu)
    e = lambda a, r: float((a.double() - r).abs().max() / (r.abs().max() + 1e-30))
    print("%-26s tag %-34s shape %-20s dx err %.3g  dgamma err %.3g  dbeta err %.3g  |dy mean|/rms %.3g" % (
        n, c["tag"], tuple(xx.shape), e(c["dx"], dx) if "dx" in c else -1, e(m.weight.grad, dg), e(m.bias.grad, db),
        float(dy.mean().abs() / dy.pow(2).mean().sqrt())))
run.close()
