"""CPU restatement (numpy) of threshold + keep-mask generation -- TEST INFRASTRUCTURE ONLY.

Follows pruners/dcfp_pruner.py:43-66 (get_thresh) and :68-92 (gen_channel_mask):

  * group g collects the scores of every scored layer with layer_group == g, in layer order;
    thresh[g] = sorted_ascending[int(size_g * global_percent)]; an empty group keeps thresh 0;
  * mask = score > thresh[group]  (strict, fp32);
  * min_keep = max(int(C * layer_keep), 1); when fewer than min_keep channels survive, the
    min_keep highest-scoring channels are switched on as well.  The reference takes them from an
    unstable torch.sort(descending=True), so ties that straddle the cut are implementation-defined;
    this restatement (and the CUDA kernel) break them lowest-index-first.
Pinned against the unmodified reference by tests/golden/sweep_c{1..4}.npz and prune_c{1..4}.npz.
"""
import numpy as np


def thresh_index(size, global_percent):
    return int(size * global_percent)


def thresholds(scores, groups, global_percent):
    out = [np.float32(0), np.float32(0)]
    for g in (0, 1):
        parts = [np.asarray(s, dtype=np.float32) for s, gg in zip(scores, groups) if gg == g]
        if not parts:
            continue
        allv = np.concatenate(parts)
        if allv.size:
            out[g] = np.sort(allv)[thresh_index(allv.size, global_percent)]
    return out


def min_keep_of(channels, layer_keep):
    k = int(channels * layer_keep)
    return k if k > 0 else 1


def masks(scores, groups, thresh, layer_keep):
    out = []
    for s, g in zip(scores, groups):
        s = np.asarray(s, dtype=np.float32)
        m = (s > np.float32(thresh[g])).astype(np.float32)
        mk = min_keep_of(s.shape[0], layer_keep)
        if int(m.sum()) < mk:
            order = np.argsort(-s, kind="stable")  # descending, ties lowest index first
            m[order[:mk]] = 1.0
        out.append(m)
    return out
