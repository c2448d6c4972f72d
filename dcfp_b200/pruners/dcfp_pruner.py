"""Host-side mirror of pruners/dcfp_pruner.py of the reference: `dcfp_pruning` (EIC score
accumulator, :7-26) and `DCFPPruner` (global thresholds + keep masks, :29-92).

Same names, arguments, state layout and on-disk format (`score.pth = {'eic': {bn_name: Tensor[C]}}`);
the arithmetic runs in the sm_100a kernels behind torch.ops.dcfp (no CPU fallback):

  * `dcfp_pruning.step`   ~10 tiny launches per BN layer in the reference (600-1100 per step)
                          -> ONE `dcfp_eic_update` launch over all layers, bit-exact fp32;
  * `get_thresh`          CPU torch.sort per group -> exact radix select (`dcfp_thresh_mask`);
  * `gen_channel_mask`    per-layer gt / sum / sort -> one CTA per layer in the same call.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from .channel_pruner import ChannelPruner


class dcfp_pruning():
    def __init__(self, model, r=0.99, **kwards):
        self.r = r
        self.state_dict = {'eic': {}}
        for name, m in model.named_modules():
            if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)) and name not in model.ignore_prune_layer:
                self.state_dict['eic'][name] = 0  # Python int until the first step, like the reference (:13)
        self._flat = None
        self._offsets = None
        self._names = None

    def _bind(self, layers):
        sizes = [m.weight.numel() for _, m in layers]
        device = layers[0][1].weight.device
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        flat = torch.zeros(offs[-1], dtype=torch.float32, device=device)
        fresh = []  # layers without a score yet (int 0): their first update treats eic as 0, like the reference (:13,:19-20)
        for (name, _), a, b in zip(layers, offs[:-1], offs[1:]):
            old = self.state_dict['eic'].get(name, 0)
            if torch.is_tensor(old):  # re-bind (the set of scored layers changed): keep what was accumulated
                flat[a:b].copy_(old.to(device).reshape(-1))
            else:
                fresh.append(name)
        self._flat = flat
        self._offsets = torch.tensor(offs, dtype=torch.int32).to(device)
        self._names = [n for n, _ in layers]
        for (name, _), a, b in zip(layers, offs[:-1], offs[1:]):
            self.state_dict['eic'][name] = self._flat[a:b]  # per-layer views of the one score vector
        return len(fresh) == len(layers)

    def step(self, model):
        """eic = eic*r + (flag*|grad| + !flag*eic)*(1-r), flag = grad*gamma > 0   (reference :15-20)."""
        layers = [(name, m) for name, m in model.named_modules() if name in self.state_dict['eic']]
        if not layers:
            return
        for name, m in layers:
            if m.weight.grad is None:
                raise AttributeError("'NoneType' object has no attribute 'data' (%s.weight.grad is None)" % name)
            if not m.weight.is_cuda:
                raise RuntimeError("dcfp_pruning.step: %s lives on %s; the EIC update runs on the GPU only" % (name, m.weight.device))
        ops.require_gpu()
        first = False
        if self._flat is None or self._names != [n for n, _ in layers]:
            # first = "every layer starts from the int 0 state": the kernel then skips reading eic.  A re-bind that carries
            # accumulated scores over takes the general path: zero-initialised entries give the same bits (0*r + g*(1-r)).
            first = self._bind(layers)
        with torch.no_grad():
            ops.eic_update([m.weight.grad.detach().contiguous() for _, m in layers],
                           [m.weight.detach().contiguous() for _, m in layers], self._offsets, self._flat, self.r, first)

    def get_eic(self):
        """{'eic': {name: Tensor[C] | int 0}}.  The per-layer tensors of `self.state_dict` are views of ONE flat device
        vector that `step` updates in place; the reference rebinds fresh tensors every step (:20), so what a caller got
        earlier keeps its values there.  Same behaviour here: the returned tensors are copies."""
        return {'eic': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.state_dict['eic'].items()}}

    def export_eic(self, path):
        out = {'eic': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.state_dict['eic'].items()}}
        torch.save(out, path)


class DCFPPruner(ChannelPruner):
    def __init__(self, global_percent=0.8, layer_keep=0.01, except_start_keys=['head.fc'], score_file='', **kwards):
        super(DCFPPruner, self).__init__(except_start_keys=except_start_keys)
        self.layer_keep = layer_keep
        self.global_percent = global_percent
        self.eic = torch.load(score_file, map_location='cpu')['eic']
        self._thresh = None
        self._selected = None  # (key, result) of the last K2 call: get_thresh() + gen_channel_mask() share one launch

    def get_bn_group(self, bn_layer):
        return 0 if bn_layer.startswith('backbone') else 1

    def get_para_score(self, bn_layer):
        return self.eic[bn_layer]

    def _select(self):
        """One K2 call: thresholds over the BNs outside `except_layers` (:43-66) and masks for every
        link whose CONV is outside `except_layers` (:68-92).  Returns (thresh[2], {bn: mask})."""
        key = (self.global_percent, self.layer_keep, id(self.eic), tuple(self.norm_conv_links.items()), tuple(sorted(self.except_layers)))
        if self._selected is not None and self._selected[0] == key:
            return self._selected[1]
        ops.require_gpu()
        device = ops.device()
        if os.environ.get("DCFP_TRACE_BACKEND"):
            print("dcfp backend: cuda (%s), torch.ops.dcfp abi v%d" % (torch.cuda.get_device_name(device), int(ops.load().abi_version())), flush=True)
        layers, scores, groups, min_keep = [], [], [], []
        bn_size = [0, 0]
        for bn_layer, conv_layer in self.norm_conv_links.items():
            in_thresh = bn_layer not in self.except_layers
            needs_mask = conv_layer not in self.except_layers
            if not (in_thresh or needs_mask):
                continue
            channels = self.name2module[bn_layer].weight.data.shape[0]
            score = self.get_para_score(bn_layer)
            group = self.get_bn_group(bn_layer)
            if in_thresh:
                bn_size[group] += channels
            layers.append((bn_layer, needs_mask))
            scores.append(torch.as_tensor(score, dtype=torch.float32).reshape(-1).cpu())
            assert scores[-1].numel() == channels, "score of %s has %d entries for %d channels" % (bn_layer, scores[-1].numel(), channels)
            groups.append(group if in_thresh else group + 2)
            keep = int(channels * self.layer_keep)
            min_keep.append(keep if keep > 0 else 1)
        if not layers:
            return [0, 0], {}
        k_idx = [int(bn_size[g] * self.global_percent) if bn_size[g] > 0 else -1 for g in (0, 1)]
        for g in (0, 1):
            if k_idx[g] >= bn_size[g] > 0:
                raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (k_idx[g], bn_size[g]))
        offs = [0]
        for s in scores:
            offs.append(offs[-1] + s.numel())
        n = len(layers)
        table = torch.tensor(offs + groups + min_keep, dtype=torch.int32).to(device)  # one host->device copy
        mask, thresh, _kept = ops.thresh_mask(torch.cat(scores).to(device), table[:n + 1], table[n + 1:2 * n + 1],
                                              table[2 * n + 1:], k_idx[0], k_idx[1])
        mask, thresh = mask.cpu(), thresh.cpu()
        out = {}
        for (bn_layer, needs_mask), a, b in zip(layers, offs[:-1], offs[1:]):
            if needs_mask:
                out[bn_layer] = mask[a:b].clone()
        res = ([thresh[g] if bn_size[g] > 0 else 0 for g in (0, 1)], out)
        self._selected = (key, res)
        return res

    def get_thresh(self):
        thresh, _ = self._select()
        return thresh

    def gen_channel_mask(self):
        thresh, masks = self._select()
        self._thresh = thresh
        for bn_layer, conv_layer in self.norm_conv_links.items():
            if conv_layer not in self.except_layers:
                conv = self.name2module[conv_layer]
                conv.out_mask = masks[bn_layer].reshape(conv.out_mask.shape).to(conv.out_mask.device)
