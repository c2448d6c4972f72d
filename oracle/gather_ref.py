"""CPU restatement (numpy) of the weight slicing and bias compensation -- TEST INFRASTRUCTURE ONLY.

Follows pruners/channel_pruner.py:907-948 (deploy_subnet):
    W = W[out_mask == 1][:, in_mask == 1];  bias / running_mean / running_var = ...[out_mask == 1]
and :873-905 (resize_subnet_bias):
    act = relu((1 - in_mask) * beta_parent);  offset = W.sum((2, 3)) @ act
Pinned against the unmodified reference by tests/golden/prune_*.npz (SHA-256 of every pruned tensor).
"""
import numpy as np


def gather(src, out_idx=None, in_idx=None):
    src = np.asarray(src)
    if out_idx is not None:
        src = src[np.asarray(out_idx, dtype=np.int64)]
    if in_idx is not None:
        src = src[:, np.asarray(in_idx, dtype=np.int64)]
    return np.ascontiguousarray(src)


def mask_to_idx(mask):
    return np.nonzero(np.asarray(mask).reshape(-1) == 1)[0].astype(np.int32)


def bias_offset(W, act):
    """fp64 evaluation of offset[o] = sum_i act[i] * sum_{kh,kw} W[o,i,kh,kw]."""
    W = np.asarray(W, dtype=np.float64)
    conv_sum = W.reshape(W.shape[0], W.shape[1], -1).sum(axis=2)
    return conv_sum @ np.asarray(act, dtype=np.float64).reshape(-1)
