"""Per-parameter gradient error of (a) the fused-BN scoring pass and (b) the cuDNN-BN + hook pass against an fp64 ARBITER
(the same model in double precision, plain autograd) on the same micro-batch, in network order, next to the run-to-run
noise of the unfused pass.  Errors are max |a - b| / max |b| per parameter tensor."""
import copy
import sys

import torch

sys.path.insert(0, ".")
from dcfp_b200.scorer import CalibrationRun  # noqa: E402
from dcfp_b200.workloads.segnets import build_segnet  # noqa: E402
from dcfp_b200.workloads.synthetic import synthetic_batch  # noqa: E402

DEV = "cuda"
K, H, W = 19, 128, 256
torch.backends.cudnn.allow_tf32 = False
model = build_segnet("deeplabv3", "resnet50", K, seed=0).to(DEV).to(memory_format=torch.channels_last)
for m in model.modules():
    if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
        m.p = 0.0
x, y = synthetic_batch([0, 1], K, H, W)
x, y = x.to(DEV).contiguous(memory_format=torch.channels_last), y.to(DEV)


def grads(fused, steps=2):
    run = CalibrationRun(model, K, seed=5, fused=fused)
    for s in range(steps):
        loss = run.step(x, y, mb_index=0)
    g = {n: p.grad.detach().double().clone() for n, p in model.named_parameters() if p.grad is not None}
    run.close()
    return float(loss), g


arb = copy.deepcopy(model).double().train()
out = arb(x.double(), y.long(), deepsup=True)
la = out["loss"] if isinstance(out, dict) else out
la.backward()
ga = {n: p.grad.detach().clone() for n, p in arb.named_parameters() if p.grad is not None}
model.train()
model.zero_grad(set_to_none=True)
out = model(x, y.long(), deepsup=True)
(out["loss"] if isinstance(out, dict) else out).backward()
gp = {n: p.grad.detach().double().clone() for n, p in model.named_parameters() if p.grad is not None}
model.zero_grad(set_to_none=True)
l0, g0 = grads(False)
l1, g1 = grads(True)
print("loss fp64 %.8f plain %.8f unfused %.8f fused %.8f" % (float(la), float(out["loss"] if isinstance(out, dict) else out), l0, l1))
print("%-40s %10s %10s %10s" % ("parameter", "fused", "unfused", "plain fp32"))
for n in ga:
    ref = ga[n].abs().max() + 1e-30
    e = [float((g[n] - ga[n]).abs().max() / ref) for g in (g1, g0, gp)]
    if n.endswith("weight"):
        print("%-40s %10.3g %10.3g %10.3g" % (n, *e))
