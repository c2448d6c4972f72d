"""CPU stand-in for torch.ops.dcfp.* built from the oracle -- TEST-ONLY.

Lets the `-m "not gpu"` suite exercise the HOST logic of the product (graph analysis, mask
propagation, gather bookkeeping, drop-in CLI flow) in a container without a GPU.  It is never
imported by the package: the product path has no CPU fallback.
"""
import contextlib

import numpy as np
import torch

from dcfp_b200 import ops
from oracle import eic_ref, gather_ref


def _thresh_mask(score, layer_off, layer_group, min_keep, k0, k1):
    s = score.cpu().numpy()
    off = layer_off.cpu().numpy()
    grp = layer_group.cpu().numpy()
    mk = min_keep.cpu().numpy()
    layers = [s[a:b] for a, b in zip(off[:-1], off[1:])]
    thresh = np.zeros(2, dtype=np.float32)
    for g, k in ((0, k0), (1, k1)):
        if k >= 0:
            allv = np.concatenate([l for l, gg in zip(layers, grp) if gg == g])
            thresh[g] = np.sort(allv)[k]
    masks, kept = [], []
    for l, g, m in zip(layers, grp, mk):
        mask = (l > thresh[g & 1]).astype(np.float32)
        if int(mask.sum()) < m:
            mask[np.argsort(-l, kind="stable")[:m]] = 1.0
        masks.append(mask)
        kept.append(int(mask.sum()))
    return torch.from_numpy(np.concatenate(masks)), torch.from_numpy(thresh), torch.tensor(kept, dtype=torch.int32)


def _gather_grouped(srcs, out_idx, in_idx):
    outs = []
    for s, o, i in zip(srcs, out_idx, in_idx):
        outs.append(torch.from_numpy(gather_ref.gather(s.cpu().numpy(), None if o is None else o.cpu().numpy(),
                                                       None if i is None else i.cpu().numpy())))
    return outs


def _bias_comp(W, act):
    return torch.from_numpy(gather_ref.bias_offset(W.cpu().numpy(), act.cpu().numpy()).astype(np.float32))


def _eic_update(grads, gammas, offsets, eic, r, first_step):
    off = offsets.cpu().numpy()
    for g, w, a, b in zip(grads, gammas, off[:-1], off[1:]):
        prev = 0 if first_step else eic[a:b].cpu().numpy()
        eic[a:b] = torch.from_numpy(eic_ref.eic_step(prev, g.cpu().numpy(), w.cpu().numpy(), r))


@contextlib.contextmanager
def oracle_backend():
    names = ["require_gpu", "device", "thresh_mask", "channel_gather_grouped", "bias_comp", "eic_update"]
    saved = {n: getattr(ops, n) for n in names}
    ops.require_gpu = lambda: None
    ops.device = lambda: torch.device("cpu")
    ops.thresh_mask = _thresh_mask
    ops.channel_gather_grouped = _gather_grouped
    ops.bias_comp = _bias_comp
    ops.eic_update = _eic_update
    try:
        yield
    finally:
        for n, v in saved.items():
            setattr(ops, n, v)
