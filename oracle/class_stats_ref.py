"""CPU restatement of K1 `class_stats` -- TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Spec source: SURVEY.md section 8 (a0).  The reference has NO counterpart for the forward functor
("parity unpinned"); the backward functor is pinned through `sum_k S1[k, c] == bn.weight.grad`,
the dgamma reduction of autograd's BN backward that pruners/dcfp_pruner.py:18 consumes.

  lab = F.interpolate(label[:, None].float(), (h, w), mode='nearest').long()   (legacy nearest)
  drop lab outside [0, K)   (ignore label 255)
  S1[k, c] = sum_{p: lab(p) = k} v[p, c],  S2[k, c] = sum v^2,  cnt[k] = #pixels       (fp64)
"""
import numpy as np
import torch


def nearest_labels(label, h, w):
    """Index-math restatement of legacy `nearest`: src = min(floor(dst * (float32)(in/out)), in-1).

    Follows ATen's nearest_neighbor_compute_source_index (UpSample.h) -- the scale and the product
    are float32.  tests/test_oracle_golden.py::test_nearest_labels_matches_interpolate checks it against F.interpolate itself.
    """
    label = torch.as_tensor(label)
    n, h0, w0 = label.shape
    sh = np.float32(h0) / np.float32(h)
    sw = np.float32(w0) / np.float32(w)
    ii = np.minimum(np.floor(np.arange(h, dtype=np.float32) * sh).astype(np.int64), h0 - 1)
    jj = np.minimum(np.floor(np.arange(w, dtype=np.float32) * sw).astype(np.int64), w0 - 1)
    return label[:, torch.from_numpy(ii)][:, :, torch.from_numpy(jj)].long()


def class_stats(v, label, K):
    """v: [N,C,h,w] (any float dtype, values are taken as given); label: [N,H0,W0] ints or None."""
    v = torch.as_tensor(v).double()
    n, c, h, w = v.shape
    if label is None:
        lab = torch.zeros(n, h, w, dtype=torch.long)
    else:
        lab = nearest_labels(label, h, w)
    flat = v.permute(0, 2, 3, 1).reshape(-1, c)
    lab = lab.reshape(-1)
    keep = (lab >= 0) & (lab < K)
    flat, lab = flat[keep], lab[keep]
    S1 = torch.zeros(K, c, dtype=torch.float64).index_add_(0, lab, flat)
    S2 = torch.zeros(K, c, dtype=torch.float64).index_add_(0, lab, flat * flat)
    cnt = torch.bincount(lab, minlength=K).double()
    return cnt, S1, S2


def functor_fwd(x, scale=None, shift=None):
    """v = x * scale[c] + shift[c], evaluated in fp64 from the stored (fp32 / bf16) inputs."""
    v = torch.as_tensor(x).double()
    if scale is not None:
        v = v * torch.as_tensor(scale).double().view(1, -1, 1, 1)
    if shift is not None:
        v = v + torch.as_tensor(shift).double().view(1, -1, 1, 1)
    return v


def functor_bwd(x, dy, scale, shift):
    """v = dy * xhat with xhat = x * invstd - mean * invstd (scale = invstd, shift = -mean*invstd)."""
    return torch.as_tensor(dy).double() * functor_fwd(x, scale, shift)


def class_stats_fwd(x, label, K, scale=None, shift=None):
    return class_stats(functor_fwd(x, scale, shift), label, K)


def class_stats_bwd(x, dy, mean, invstd, label, K):
    scale = torch.as_tensor(invstd).float()
    shift = (-torch.as_tensor(mean).float() * scale)
    return class_stats(functor_bwd(x, dy, scale, shift), label, K)


def abs_mass(v, label, K):
    """sum |v| per (class, channel): the scale against which fp32-accumulation error is judged."""
    return class_stats(torch.as_tensor(v).double().abs(), label, K)[1]


def outside_stats(v, label, K):
    """(S1[C], S2[C]) over the pixels whose nearest-down-sampled label is OUTSIDE [0, K) -- the ignore label.  Those
    pixels carry no loss (criterion.py:52-60, ignore_index) but they do carry gradient at every layer below the logits,
    and autograd's bn.weight.grad sums over them: sum_k S1[k] + outside S1 == dgamma, not sum_k S1[k] alone.  The
    scorer keeps them in row K of its arenas (dcfp_b200/scorer.py)."""
    v = torch.as_tensor(v).double()
    n, c, h, w = v.shape
    lab = nearest_labels(label, h, w).reshape(-1)
    flat = v.permute(0, 2, 3, 1).reshape(-1, c)
    out = flat[(lab < 0) | (lab >= K)]
    return out.sum(0), (out * out).sum(0)
