"""On-disk formats either side of the path: `score.pth`, `channel_cfg.pth`, `pruned.pth`.

The reference has no format module; the three files are written and read inline:

  score.pth        torch.save({'eic': {bn_name: Tensor[C] | int 0}})       pruners/dcfp_pruner.py:25-26, train.py:287
                   read by DCFPPruner.__init__                              pruners/dcfp_pruner.py:34
  channel_cfg.pth  torch.save({module_name: {in_channels, raw_in_channels, in_mask: np[1,C,1,1] f32, out_...}})
                                                                            prune.py:98, pruners/channel_pruner.py:821-842
                   read by prune.py:108, train.py:201, evaluate.py, totrt.py (plain torch.load)
  pruned.pth       torch.save(sub_model.state_dict())                       prune.py:97
                   read by utils/pyt_utils.py:43-96 (load_model) after init_pruned_model

Files written here are readable by the unmodified reference, and files the reference wrote are readable
here.  Two facts a caller on torch >= 2.6 trips over are handled in one place:

  * `torch.load` defaults to `weights_only=True`, which rejects the numpy masks inside a reference-written
    `channel_cfg.pth` (prune.py:108 fails there).  `load_channel_cfg` allow-lists exactly the numpy
    reconstruction globals such a file needs -- it does not fall back to unrestricted unpickling.
  * `save_channel_cfg(..., portable=True)` stores the masks as tensors, so any consumer can read the file
    with the default `torch.load`; `load_channel_cfg` turns them back into the numpy arrays `export_subnet`
    produces, so both spellings compare equal after loading.

Host-side only (file I/O and shape bookkeeping): nothing here launches a kernel.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from .channel_pruner import init_pruned_model

__all__ = ["save_score", "load_score", "save_channel_cfg", "load_channel_cfg", "save_pruned", "load_state",
           "load_pruned_model"]

_MASK_KEYS = ("in_mask", "out_mask")


# ---- score.pth -----------------------------------------------------------------------------------------------

def save_score(eic, path):
    """Write `{'eic': {name: Tensor[C]}}` exactly as `dcfp_pruning.export_eic` does (dcfp_pruner.py:25-26).

    Tensors go to the CPU first: the reference saves CUDA tensors and reads them back with
    `map_location='cpu'` (:34); a CPU file loads on a box without a GPU either way.  A layer that never took
    a step keeps the reference's initial state, the Python int 0 (:13)."""
    out = OrderedDict()
    for name, v in eic.items():
        out[name] = v.detach().to("cpu", copy=True).contiguous() if torch.is_tensor(v) else v
    torch.save({"eic": out}, path)


def load_score(path):
    """`torch.load(score_file, map_location='cpu')['eic']` (dcfp_pruner.py:34); tensors and ints only, so the
    restricted unpickler is enough."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(blob, dict) or "eic" not in blob:
        raise KeyError("'eic'")  # the reference's own failure for a file that is not a score file
    return blob["eic"]


# ---- channel_cfg.pth -----------------------------------------------------------------------------------------

def _numpy_safe_globals():
    """The globals pickle needs to rebuild a float32 ndarray, and nothing else."""
    out = [np.ndarray, np.dtype, type(np.dtype(np.float32))]
    try:  # numpy >= 2
        from numpy._core.multiarray import _reconstruct
    except ImportError:  # numpy 1.x wrote numpy.core.multiarray
        from numpy.core.multiarray import _reconstruct
    out.append(_reconstruct)
    return out


def save_channel_cfg(channel_cfg, path, portable=False):
    """prune.py:98.  `portable=False` writes the reference's exact layout (numpy masks); `portable=True` writes
    the masks as float32 tensors of the same shape so the default restricted `torch.load` reads the file."""
    if not portable:
        torch.save(channel_cfg, path)
        return
    out = OrderedDict()
    for name, cfg in channel_cfg.items():
        row = dict(cfg)
        for k in _MASK_KEYS:
            if k in row and isinstance(row[k], np.ndarray):
                row[k] = torch.from_numpy(np.ascontiguousarray(row[k]))
        out[name] = row
    torch.save(out, path)


def load_channel_cfg(path):
    """Read a `channel_cfg.pth` written by the reference (numpy masks) or by `save_channel_cfg(portable=True)`;
    returns the `export_subnet` layout (`channel_pruner.py:821-842`): int counts + numpy `[1,C,1,1]` f32 masks."""
    with torch.serialization.safe_globals(_numpy_safe_globals()):
        blob = torch.load(path, map_location="cpu", weights_only=True)
    out = OrderedDict()
    for name, cfg in blob.items():
        row = dict(cfg)
        for k in _MASK_KEYS:
            if k in row and torch.is_tensor(row[k]):
                row[k] = row[k].numpy()
        for k in ("in_channels", "out_channels", "raw_in_channels", "raw_out_channels"):
            if k in row:
                row[k] = int(row[k])
        out[name] = row
    return out


# ---- pruned.pth ----------------------------------------------------------------------------------------------

def save_pruned(sub_model, channel_cfg, save_path, portable=False):
    """The two writes of prune.py:97-98 into directory `save_path`; returns the two file names."""
    os.makedirs(save_path, exist_ok=True)
    weights = os.path.join(save_path, "pruned.pth")
    cfg = os.path.join(save_path, "channel_cfg.pth")
    torch.save(OrderedDict((k, v.detach().to("cpu")) for k, v in sub_model.state_dict().items()), weights)
    save_channel_cfg(channel_cfg, cfg, portable=portable)
    return weights, cfg


def load_state(model, model_file, ignore_prefix=None, extra_prefix=None):
    """`utils/pyt_utils.py:43-96` (load_model) without its logger: accept a path or a state dict, unwrap a
    `'model'` / `'state_dict'` envelope, strip `ignore_prefix` / prepend `extra_prefix`, load non-strictly and
    return `(missing, unexpected)` key lists with `.num_batches_tracked` entries left out, as the reference
    reports them.  (`extra_prefix` raises NameError in the reference, :66 -- here it does what :63-68 intends.)"""
    if isinstance(model_file, (str, os.PathLike)):
        state = torch.load(model_file, map_location="cpu", weights_only=True)
        for envelope in ("model", "state_dict"):
            if envelope in state:
                state = state[envelope]
                break
    else:
        state = model_file
    if ignore_prefix is not None:
        state = OrderedDict((k[len(ignore_prefix):] if k.startswith(ignore_prefix) else k, v)
                            for k, v in state.items())
    if extra_prefix is not None:
        state = OrderedDict((extra_prefix + k, v) for k, v in state.items())
    model.load_state_dict(state, strict=False)
    own = set(model.state_dict().keys())
    got = set(state.keys())
    tracked = ".num_batches_tracked"
    missing = sorted(k for k in own - got if not k.endswith(tracked))
    unexpected = sorted(k for k in got - own if not k.endswith(tracked))
    return missing, unexpected


def load_pruned_model(model, channel_cfg, weights, strict=True):
    """prune.py:108-110 / train.py:200-207 / evaluate.py:289: slice a FRESH `model` to the sizes in `channel_cfg`
    (a dict or a path), then load `weights` (a state dict or a path).  With `strict`, a key that is missing or
    left over raises instead of being logged -- after `init_pruned_model` every shape must line up."""
    if isinstance(channel_cfg, (str, os.PathLike)):
        channel_cfg = load_channel_cfg(channel_cfg)
    init_pruned_model(model, channel_cfg)
    missing, unexpected = load_state(model, weights)
    if strict and (missing or unexpected):
        raise RuntimeError("pruned weights do not match the model: missing %s, unexpected %s" % (missing, unexpected))
    return model
