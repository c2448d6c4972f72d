"""Threshold margins: how far the scores sit from the group thresholds that decide the keep masks.

`DCFPPruner.gen_channel_mask` keeps a channel iff `score > thresh[group]` (strict, pruners/dcfp_pruner.py:77) with
`thresh[group]` = the int(size * global_percent)-th smallest score of the group (:59-64) -- an element of the score vector.
Two scoring stacks (1 vs N GPUs, fused vs unfused BN, CPU reference vs GPU) agree on the scores to a tolerance only, so
identical masks are guaranteed only when no score lies within that tolerance of its threshold.  These helpers put a number
on it (SURVEY.md section 7.3 step 4, section 8 e): every cross-stack mask comparison reports the margin next to the verdict.
Host-side numpy on the small score vectors (tens of thousands of floats): bookkeeping, not a kernel.
"""
import numpy as np


def group_values(score, layer_off, layer_group):
    score = np.asarray(score, dtype=np.float32)
    out = {}
    for g in sorted(set(int(x) & 1 for x in layer_group)):
        idx = np.concatenate([np.arange(layer_off[i], layer_off[i + 1]) for i, gg in enumerate(layer_group) if int(gg) == g] or
                             [np.zeros(0, dtype=np.int64)])
        out[g] = (idx, score[idx])
    return out


def threshold_margins(score, layer_off, layer_group, global_percent):
    """{group: dict(thresh, below, above, margin)}: `below` / `above` = the nearest scores strictly below / above the
    threshold, `margin` = min(thresh - below, above - thresh) / thresh -- the relative perturbation of a single score that
    can flip a mask bit (ties AT the threshold are pruned together by the strict compare and do not count)."""
    res = {}
    for g, (idx, vals) in group_values(score, layer_off, layer_group).items():
        if vals.size == 0:
            continue
        s = np.sort(vals)
        t = s[min(int(vals.size * global_percent), vals.size - 1)]
        lo = s[s < t]
        hi = s[s > t]
        below = float(lo[-1]) if lo.size else float("nan")
        above = float(hi[0]) if hi.size else float("nan")
        gaps = [x for x in (float(t) - below, above - float(t)) if x == x]
        # thresh == 0 (after few steps half of the scores are exactly 0: the gate failed every time) has no relative margin:
        # every zero score is pruned together, the margin is the absolute gap to the smallest positive score
        rel = float(t) != 0.0
        res[g] = dict(thresh=float(t), below=below, above=above, relative=rel,
                      margin=((min(gaps) / abs(float(t))) if rel else min(gaps)) if gaps else float("inf"))
    return res


def compare_masks(score_a, score_b, layer_off, layer_group, global_percent):
    """Masks of two score vectors for the same layers (thresholds per vector; the min-keep fallback is left out: it only
    adds channels).  Returns dict(flipped, n, flip_band, disc_near, margins_a, margins_b):
      flipped    channels whose keep bit differs
      flip_band  max over the flipped channels of min_{a,b} |score - thresh| / thresh (0 if none): a channel whose keep bit
                 differs sits, in at least one of the two vectors, this close to its threshold -- it must be inside the
                 band the score discrepancy explains
      disc_near  max relative discrepancy |a - b| / max(a, b) over the channels within 10 % of a threshold."""
    a, b = np.asarray(score_a, dtype=np.float32), np.asarray(score_b, dtype=np.float32)
    ma, mb = threshold_margins(a, layer_off, layer_group, global_percent), threshold_margins(b, layer_off, layer_group, global_percent)
    flipped, band, disc = 0, 0.0, 0.0
    for g, (idx, va) in group_values(a, layer_off, layer_group).items():
        vb = b[idx]
        ta, tb = ma[g]["thresh"], mb[g]["thresh"]
        ka, kb = va > ta, vb > tb
        da = np.abs(va - ta) / max(abs(ta), 1e-38)
        db = np.abs(vb - tb) / max(abs(tb), 1e-38)
        f = ka != kb
        flipped += int(f.sum())
        if f.any():
            band = max(band, float(np.minimum(da[f], db[f]).max()))
        near = (da < 0.1) | (db < 0.1)
        if near.any():
            disc = max(disc, float((np.abs(va - vb)[near] / np.maximum(np.maximum(va, vb)[near], 1e-38)).max()))
    return dict(flipped=flipped, n=int(a.size), flip_band=band, disc_near=disc, margins_a=ma, margins_b=mb)


def format_margins(m):
    return ", ".join("group %d: thresh %.6g, margin %.3g%s" % (g, v["thresh"], v["margin"], "" if v.get("relative", True) else " (absolute: thresh is 0)")
                     for g, v in sorted(m.items()))
