"""Shape-only model cost: the numbers `utils/flops_counter.get_model_complexity_info` reports (multiply-add = 1 FLOP,
utils/flops_counter.py:115), obtained from a forward pass on the META device.

The reference runs a real CPU forward of the network on an uninitialised 512x512 batch (8-15 s per call,
flops_counter.py:85-99) once before and once inside every iteration of prune.py's search loop (prune.py:78,112) just to
read tensor SHAPES in forward hooks.  Here the same hooks see meta tensors: no arithmetic, a few tens of milliseconds.

Per-module cost, as the reference counts it (flops_counter.py:372-459; only `type(module)`-exact matches count):
  Conv1d/2d/3d       out_positions * (prod(kernel) * in_channels * out_channels / groups [+ out_channels if bias])
  ReLU family        output elements (every call of a shared module counts)
  Max/Avg/Adaptive   input elements
  Batch/Group/Instance/LayerNorm   input elements, x2 when affine
  Linear             input elements * out_features
  Upsample (module)  output elements            (F.interpolate calls are NOT modules and cost nothing)
"""
import copy

import numpy as np
import torch
import torch.nn as nn


def _conv(m, inp, out):
    x = inp[0]
    per_position = int(np.prod(m.kernel_size)) * m.in_channels * (m.out_channels // m.groups)
    active = x.shape[0] * int(np.prod(out.shape[2:]))
    return per_position * active + (m.out_channels * active if m.bias is not None else 0)


def _out_numel(m, inp, out):
    return out.numel()


def _in_numel(m, inp, out):
    return int(np.prod(inp[0].shape))


def _norm(m, inp, out):
    n = int(np.prod(inp[0].shape))
    return 2 * n if (getattr(m, "affine", False) or getattr(m, "elementwise_affine", False)) else n


def _linear(m, inp, out):
    return int(np.prod(inp[0].shape) * out.shape[-1])


def _upsample(m, inp, out):
    return int(np.prod(out[0].shape))  # the reference indexes `output[0]`: the first sample only (flops_counter.py:373)


_COST = {}
for _t in (nn.Conv1d, nn.Conv2d, nn.Conv3d):
    _COST[_t] = _conv
for _t in (nn.ReLU, nn.PReLU, nn.ELU, nn.LeakyReLU, nn.ReLU6):
    _COST[_t] = _out_numel
for _t in (nn.MaxPool1d, nn.AvgPool1d, nn.AvgPool2d, nn.MaxPool2d, nn.MaxPool3d, nn.AvgPool3d, nn.AdaptiveMaxPool1d,
           nn.AdaptiveAvgPool1d, nn.AdaptiveMaxPool2d, nn.AdaptiveAvgPool2d, nn.AdaptiveMaxPool3d, nn.AdaptiveAvgPool3d):
    _COST[_t] = _in_numel
for _t in (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.GroupNorm, nn.InstanceNorm1d, nn.InstanceNorm2d,
           nn.InstanceNorm3d, nn.LayerNorm):
    _COST[_t] = _norm
_COST[nn.Linear] = _linear
_COST[nn.Upsample] = _upsample


def to_meta(model):
    """A structural copy of `model` whose parameters / buffers live on the meta device (no storage)."""
    memo = {id(t): torch.empty_like(t, device="meta") if not isinstance(t, nn.Parameter)
            else nn.Parameter(torch.empty_like(t, device="meta"), requires_grad=t.requires_grad)
            for t in list(model.parameters()) + list(model.buffers())}
    return copy.deepcopy(model, memo)


def model_cost(model, input_shape=(3, 512, 512)):
    """(flops, params) as plain numbers, like get_model_complexity_info(..., as_strings=False)."""
    meta = model if next(model.parameters()).device.type == "meta" else to_meta(model)
    meta.eval()
    total = [0]
    handles = []
    for m in meta.modules():
        fn = _COST.get(type(m))
        if fn is not None:
            handles.append(m.register_forward_hook(lambda mod, i, o, fn=fn: total.__setitem__(0, total[0] + int(fn(mod, i, o)))))
    try:
        with torch.no_grad():
            meta(torch.empty((1,) + tuple(input_shape), dtype=next(meta.parameters()).dtype, device="meta"))
    finally:
        for h in handles:
            h.remove()
    params = sum(p.numel() for p in meta.parameters() if p.requires_grad)
    return total[0] / 1, params


def flops_to_string(flops, precision=2):
    """'<x> GFLOPs' exactly as the reference prints it (flops_counter.py:143): prune.py parses this string."""
    return str(round(flops / 10. ** 9, precision)) + " GFLOPs"


def params_to_string(num_params, precision=2):
    if num_params // 10 ** 6 > 0:
        return str(round(num_params / 10 ** 6, precision)) + " M"
    if num_params // 10 ** 3:
        return str(round(num_params / 10 ** 3, precision)) + " k"
    return str(num_params)


def get_model_complexity_info(model, input_shape, print_per_layer_stat=False, as_strings=True, **_):
    """Drop-in for utils.flops_counter.get_model_complexity_info as prune.py calls it (prune.py:78,112)."""
    flops, params = model_cost(model, input_shape)
    if as_strings:
        return flops_to_string(flops), params_to_string(params)
    return flops, params
