"""Host-side mirror of pruners/random_pruner.py (random baseline the reference's prune.py imports, :13)."""
import torch

from .channel_pruner import ChannelPruner


class RandomChannelPruner(ChannelPruner):
    def __init__(self, global_percent=0.8, layer_keep=0.01, except_start_keys=['head.fc'], **kwards):
        super(RandomChannelPruner, self).__init__(except_start_keys=except_start_keys)
        self.layer_keep = layer_keep
        self.global_percent = global_percent

    def gen_channel_mask(self):
        """mask = rand(C) > global_percent from the global torch RNG; the first `min_keep` channels are
        switched on when too few survive (reference :11-33)."""
        for bn_layer, conv_layer in self.norm_conv_links.items():
            if conv_layer in self.except_layers:
                continue
            channels = self.name2module[bn_layer].weight.shape[0]
            keep = int(channels * self.layer_keep)
            min_channel_num = keep if keep > 0 else 1
            mask = (torch.rand(channels) > self.global_percent) * 1.0
            if int(torch.sum(mask)) < min_channel_num:
                mask[:min_channel_num] = 1.
            conv = self.name2module[conv_layer]
            conv.out_mask = mask.reshape(conv.out_mask.shape)
