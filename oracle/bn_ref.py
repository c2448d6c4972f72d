"""Training-mode BatchNorm2d (+ ReLU) forward / backward in the FACTORED form a fused BN + K1 kernel pair computes --
TEST INFRASTRUCTURE ONLY (not used by any product path yet).

The reference uses plain `nn.BatchNorm2d` followed by a shared in-place `nn.ReLU` (networks/backbone/resnet.py:26-31,
networks/aspp.py:15-16); autograd's BN backward is the producer of the `bn.weight.grad` that pruners/dcfp_pruner.py:18
reads.  Everything a fused implementation needs from a feature map are per-channel SUMS -- exactly what K1 reduces:

  forward    n = N*h*w,  mean = S1x / n,  var = S2x / n - mean^2,  invstd = rsqrt(var + eps)      (S1x, S2x: K1 forward
             y = relu(gamma * (x - mean) * invstd + beta)                                          functor, one class)
  backward   gate = (gamma * xhat + beta > 0)            recomputed from x, no mask tensor
             dz   = dy * gate
             dbeta  = sum dz,   dgamma = sum dz * xhat     (K1 backward functor, summed over its class rows + outside row)
             dx = gamma * invstd * (dz - dbeta / n - xhat * dgamma / n)

`tests/test_oracle_golden.py::test_factored_batchnorm_equals_autograd` checks this against torch autograd in fp64.
"""
import torch


def bn_relu_forward(x, gamma, beta, eps=1e-5, relu=True):
    """-> (y, mean, invstd); x [N,C,h,w].  Batch statistics from the two sums K1's forward functor produces."""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    S1 = x.sum(dim=(0, 2, 3))
    S2 = (x * x).sum(dim=(0, 2, 3))
    mean = S1 / n
    var = S2 / n - mean * mean  # biased variance, as F.batch_norm normalises with
    invstd = torch.rsqrt(var + eps)
    c = (1, -1, 1, 1)
    y = (x - mean.view(c)) * (invstd * gamma).view(c) + beta.view(c)
    return (torch.relu(y) if relu else y), mean, invstd


def bn_relu_backward(x, dy, gamma, beta, mean, invstd, relu=True):
    """-> (dx, dgamma, dbeta) from x, the incoming gradient and the saved batch statistics."""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    c = (1, -1, 1, 1)
    xhat = (x - mean.view(c)) * invstd.view(c)
    dz = dy * ((gamma.view(c) * xhat + beta.view(c)) > 0).to(dy.dtype) if relu else dy
    dbeta = dz.sum(dim=(0, 2, 3))
    dgamma = (dz * xhat).sum(dim=(0, 2, 3))
    dx = (gamma * invstd).view(c) * (dz - dbeta.view(c) / n - xhat * dgamma.view(c) / n)
    return dx, dgamma, dbeta


def running_stats_update(running_mean, running_var, mean, var_biased, n, momentum=0.1):
    """What F.batch_norm does to the buffers in training mode: the running variance takes the UNBIASED estimate."""
    unbiased = var_biased * (n / max(n - 1, 1))
    return (1 - momentum) * running_mean + momentum * mean, (1 - momentum) * running_var + momentum * unbiased
