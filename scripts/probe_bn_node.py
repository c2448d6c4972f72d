"""What autograd node / saved tensors does CUDA BatchNorm produce on this box? (development aid)"""
import torch
bn = torch.nn.BatchNorm2d(8).cuda()
x = torch.randn(2, 8, 4, 4, device="cuda", requires_grad=True)
y = bn(x)
n = y.grad_fn
print(type(n).__name__, [a for a in dir(n) if a.startswith("_saved")])
for a in ("_saved_result1", "_saved_result2"):
    t = getattr(n, a, None)
    print(a, None if t is None else (t.shape, t.dtype))
m = x.detach().mean((0, 2, 3)); v = x.detach().var((0, 2, 3), unbiased=False)
print("mean ok", torch.allclose(n._saved_result1, m, atol=1e-6), "invstd ok", torch.allclose(n._saved_result2, torch.rsqrt(v + bn.eps), atol=1e-4))
