"""Random-score baseline behind the same pruner interface (the reference's prune.py imports it, prune.py:13).

Semantics kept from pruners/random_pruner.py:11-33 of the reference: one `torch.rand(C)` draw from the GLOBAL torch
generator per prunable (BN, conv) link, in link-discovery order; a channel survives when its draw exceeds
`global_percent`; a layer left with fewer than max(int(C * layer_keep), 1) survivors additionally keeps its FIRST
that-many channels.  No device work: the masks feed the same propagation / K3 gather path as DCFPPruner's.
"""
import torch

from .channel_pruner import ChannelPruner


def _floor_keep(channels, layer_keep):
    return max(int(channels * layer_keep), 1)


class RandomChannelPruner(ChannelPruner):
    def __init__(self, global_percent=0.8, layer_keep=0.01, except_start_keys=['head.fc'], **kwards):
        ChannelPruner.__init__(self, except_start_keys=except_start_keys)
        self.global_percent, self.layer_keep = global_percent, layer_keep

    def draw_mask(self, channels):
        """fp32 0/1 vector of one layer (consumes `channels` uniforms of the global generator)."""
        keep = torch.rand(channels).gt(self.global_percent).to(torch.float32)
        floor = _floor_keep(channels, self.layer_keep)
        if int(keep.sum()) < floor:
            keep[:floor] = 1.0
        return keep

    def gen_channel_mask(self):
        prunable = [(bn, conv) for bn, conv in self.norm_conv_links.items() if conv not in self.except_layers]
        for bn, conv in prunable:
            target = self.name2module[conv]
            target.out_mask = self.draw_mask(self.name2module[bn].weight.shape[0]).reshape(target.out_mask.shape)
