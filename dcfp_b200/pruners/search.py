"""prune.py's outer loop (prune.py:91-124) without its per-iteration costs.

The reference scans global_percent = 0.5, 0.52, ... and, for EVERY candidate, deep-copies the network, re-traces its
graph, slices all weights, writes two files, rebuilds a fresh model from them and runs a real 512x512 CPU forward to
count FLOPs -- ~25-30 s per candidate -- stopping at the first one whose FLOPs ratio is <= 1 - prune_ratio.  The
decision only needs channel COUNTS: here every candidate costs one K2 call (thresholds + masks, ~0.1 ms on the GPU), the
host mask propagation and a meta-device forward of a shape-only model; the weights are gathered (K3) once, for the
winner.  Same candidates (accumulated in floating point like prune.py:122), same stopping rule on the same 2-decimal
GFLOPs strings, same outputs.
"""
import copy

from . import flops as _flops
from .channel_pruner import _structural_clone, init_pruned_model
from .dcfp_pruner import DCFPPruner


def search_global_percent(model, score_file, prune_ratio=0.6, start_global_percent=0.5, step_global_percent=0.02,
                          layer_keep=0.02, except_start_keys=("conv_deepsup",), input_shape=(3, 512, 512),
                          base_flops=None, verbose=False):
    """Returns (global_percent, channel_cfg, trace) of the first candidate with flops_ratio <= 1 - prune_ratio, or of
    the last candidate below 1.0 when none qualifies (prune.py:120-124).  `model` is not modified.

    base_flops: GFLOPs of the unpruned network as prune.py measures it (built with deepsup=False, prune.py:70-79);
    defaults to this model's own cost."""
    meta_full = _flops.to_meta(model)
    if base_flops is None:
        base_flops = float(_flops.flops_to_string(_flops.model_cost(meta_full, input_shape)[0]).split(" GFLOPs")[0])
    pruner = DCFPPruner(global_percent=start_global_percent, layer_keep=layer_keep, score_file=score_file)
    clone = _structural_clone(model)
    pruner.end_nodes = getattr(clone, "end_nodes", [])
    pruner.prepare_from_supernet(clone)
    pruner.except_start_keys = pruner.except_start_keys + list(getattr(clone, "ignore_prune_layer", [])) + list(except_start_keys)
    pruner.get_except_layers(clone)
    trace = []
    global_percent = start_global_percent
    while True:
        pruner.global_percent = global_percent
        for m in pruner.name2module.values():  # fresh masks for this candidate
            for attr in ("in_mask", "out_mask"):
                if hasattr(m, attr):
                    getattr(m, attr).fill_(1.0)
        pruner.gen_channel_mask()
        pruner.set_subnet(pruner.sample_subnet())
        channel_cfg = pruner.export_subnet()
        shape_only = copy.deepcopy(meta_full)
        init_pruned_model(shape_only, channel_cfg)
        flops2 = float(_flops.flops_to_string(_flops.model_cost(shape_only, input_shape)[0]).split(" GFLOPs")[0])
        ratio = flops2 / base_flops
        trace.append((global_percent, ratio))
        if verbose:
            print("global_percent: {}, flops_ratio: {}".format(global_percent, ratio))
        if ratio <= (1 - prune_ratio):
            break
        nxt = global_percent + step_global_percent
        if nxt >= 1.0:
            break
        global_percent = nxt
    return global_percent, channel_cfg, trace


def prune_to_flops_ratio(model, score_file, prune_ratio=0.6, **kw):
    """search_global_percent + ONE prune_model at the selected percent.  Returns (sub_model, channel_cfg, global_percent);
    `model` is sliced in place like DCFPPruner.prune_model does (pruners/channel_pruner.py:967-990)."""
    layer_keep = kw.get("layer_keep", 0.02)
    except_keys = list(kw.get("except_start_keys", ("conv_deepsup",)))
    gp, _, trace = search_global_percent(model, score_file, prune_ratio, **kw)
    pruner = DCFPPruner(global_percent=gp, layer_keep=layer_keep, score_file=score_file)
    sub, channel_cfg = pruner.prune_model(model, except_start_keys=except_keys)
    return sub, channel_cfg, gp
