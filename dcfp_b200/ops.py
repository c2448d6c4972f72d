"""Loader and thin wrappers for torch.ops.dcfp.* (the C ABI registered as PyTorch ops).

There is NO fallback: if the native library is missing, or a tensor lives on the CPU, the call
raises.  `require_gpu()` is what product entry points call first.
"""
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OPS_LIB = os.path.join(HERE, "lib", "dcfp_torch_ops.so")

_loaded = False
#: DCFP_NVTX=1 wraps every call into this library in an NVTX range (dcfp::<op>) for nsys / ncu --nvtx timelines
_NVTX = os.environ.get("DCFP_NVTX", "0") == "1"


class _Nvtx:
    """Proxy over torch.ops.dcfp that brackets each op call with an NVTX range."""

    def __init__(self, ns):
        self._ns = ns

    def __getattr__(self, name):
        fn = getattr(self._ns, name)

        def wrapped(*a, **kw):
            torch.cuda.nvtx.range_push("dcfp::" + name)
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.nvtx.range_pop()

        return wrapped



# warps per CTA of the NHWC forward-functor kernel on bf16 maps (kNhwcFwdWarpsBf16 in csrc/k1_nhwc.cuh; fp32 maps: 8;
# DCFP_K1_FWD_WARPS=8|16 forces one kernel for both dtypes)
K1_FWD_WARPS_BF16 = 16


def load():
    """Registers torch.ops.dcfp.* (idempotent)."""
    global _loaded
    if _loaded:
        return _Nvtx(torch.ops.dcfp) if _NVTX else torch.ops.dcfp
    if not os.path.exists(OPS_LIB):
        raise RuntimeError("dcfp_b200 native library not built: %s is missing. Run `python -m dcfp_b200.build` "
                           "(nvcc, sm_100a). There is no CPU or eager fallback." % OPS_LIB)
    torch.ops.load_library(OPS_LIB)
    _loaded = True
    return _Nvtx(torch.ops.dcfp) if _NVTX else torch.ops.dcfp


def require_gpu():
    if not torch.cuda.is_available():
        raise RuntimeError("dcfp_b200 needs a CUDA device (B200 / sm_100a): the scoring, mask and gather kernels "
                           "have no CPU fallback.")
    return load()


def device():
    """The CUDA device product code stages tensors on."""
    return torch.device("cuda", torch.cuda.current_device())


def launch_count(reset=False):
    return int(load().launch_count(bool(reset)))


# ---- K1 ----------------------------------------------------------------------------------------
def label_keys(label, h, w, K, cnt=None):
    """uint8 class keys [N,h,w] (legacy-nearest down-sampling; labels outside [0,K) -> K = dropped);
    cnt[k] += pixels of class k (fp64, optional)."""
    return load().label_keys(label, int(h), int(w), int(K), cnt)


AFFINE_SCALE_SHIFT, AFFINE_INVSTD_MEAN = 0, 1


def class_stats(x, keys, K, S1, S2, dy=None, scale=None, shift=None, affine_mode=AFFINE_SCALE_SHIFT):
    """S1[k,c] += sum v, S2[k,c] += sum v*v over the pixels whose class key is k (keys from label_keys).
    affine_mode=AFFINE_INVSTD_MEAN: scale = invstd, shift = batch mean (what autograd's BN node saved)."""
    load().class_stats(x, dy, scale, shift, keys, S1, S2, int(K), int(affine_mode))


def class_stats_grouped(xs, keys, K, S1s, S2s, dys=None, scales=None, shifts=None, affine_mode=AFFINE_SCALE_SHIFT):
    """One launch over many resident feature maps (same dtype / K / functor); keys: one tensor per layer."""
    load().class_stats_grouped(list(xs), list(dys or []), list(scales or []), list(shifts or []), list(keys or []), list(S1s),
                               list(S2s), int(K), int(affine_mode))


def fold_step(step, total=None, step32=None):
    """dgamma[c] = sum_k step[0,k,c]; total += step; step = 0 -- one launch.  step/total: fp64 [2,K,C]; step32: optional
    fp32 arena of the same shape (the fused BN backward's class rows), folded in and zeroed the same way."""
    return load().fold_step(step, total, step32)


def reduce_classes(S1):
    return load().reduce_classes(S1)


def class_balance_weights(label, K, sample_class=None, mode=2, beta=0.9999, ignore_label=255):
    """Per-pixel class-balance weights of the finetune data path (datasets/Base.py:73-89): -> (weight fp64 [N,H,W],
    class_num int64 [N,K]).  mode 1: 1/(count+1); mode 2: effective-number ratio w.r.t. sample_class[n]."""
    return load().class_balance_weights(label, int(K), sample_class, int(mode), float(beta), int(ignore_label))


# ---- f1: fused BatchNorm2d (+ReLU) with the class-keyed sums in its backward ----------------------
def bn_supported(x):
    """True when the fused BN kernels take this feature map (channels_last fp32 / bf16, 16-byte channel vectors,
    at least 64 pixels); anything else stays with torch's BN and the hook path."""
    if x.dim() != 4 or x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda:
        return False
    if not x.is_contiguous(memory_format=torch.channels_last):
        return False
    return bool(load().bn_supported(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype == torch.bfloat16))


def bn_scratch_elems(C):
    """fp64 elements of the zeroed scratch one fused BN call (forward or backward) needs for C channels."""
    return (int(load().bn_scratch_bytes(int(C))) + 7) // 8


def bn_scratch(C, device):
    return torch.zeros(bn_scratch_elems(C), dtype=torch.float64, device=device)


def bn_workspace(C, device):
    """Unzeroed scratch the one-launch forward needs (per-CTA partial sums); one buffer of the largest C serves every
    layer of a model on one stream."""
    return torch.empty((int(load().bn_workspace_bytes(int(C))) + 3) // 4, dtype=torch.float32, device=device)


def bn_forward(x, gamma, beta, running_mean, running_var, sums, momentum, eps, relu, phases=0, workspace=None, residual=None):
    """-> (y, mean, invstd).  sums: zeroed scratch (bn_scratch).  phases: 0 = whole call; 1 = statistics pass only
    (mean, invstd, running statistics; y is empty); 2 = normalise pass only (after a phases=1 call on the same scratch).
    workspace (bn_workspace): with phases == 0 the forward is ONE cooperative launch (statistics + normalise).
    residual (same shape / dtype / layout as x): y = [relu](bn(x) + residual), the tail of a bottleneck block in one pass."""
    return load().bn_forward(x, gamma, beta, running_mean, running_var, sums, float(momentum), float(eps), bool(relu), int(phases),
                             workspace, residual)


def bn_backward(x, dy, gamma, beta, mean, invstd, keys, S1, S2, K, sums, relu, need_dx=True, phases=0):
    """-> (dx | empty, dgamma, dbeta); S1[k,c] += sum dz*xhat, S2 += (dz*xhat)^2 over the pixels of class k.
    phases: 0 = whole call; 1 = reduction pass only (sums, S1, S2); 2 = dx pass only (after a phases=1 call)."""
    return load().bn_backward(x, dy, gamma, beta, mean, invstd, keys, S1, S2, int(K), sums, bool(relu), bool(need_dx), int(phases))


def relu_grad(y, dy, dy2=None):
    """dz = (y > 0) ? dy (+ dy2) : 0 -- the ReLU backward behind a residual sum, with the accumulation of the block input's two
    gradients folded in (dy2: the shortcut gradient handed back by the next block).  Same values as torch's add + threshold_backward."""
    return load().relu_grad(y, dy, dy2)


# ---- K2 ----------------------------------------------------------------------------------------
def r_pair(r):
    """(float32(r), float32(1 - r)) exactly as `eic*r + g*(1-r)` sees them (dcfp_pruner.py:20)."""
    return float(torch.tensor(float(r), dtype=torch.float32)), float(torch.tensor(1.0 - float(r), dtype=torch.float32))


def eic_update(grads, gammas, offsets, eic, r, first_step):
    rr, omr = r_pair(r)
    load().eic_update(list(grads), list(gammas), offsets, eic, rr, omr, bool(first_step))


def eic_update_flat(grad, gamma, eic, r, first_step):
    rr, omr = r_pair(r)
    load().eic_update_flat(grad, gamma, eic, rr, omr, bool(first_step))


def thresh_mask(score, layer_off, layer_group, min_keep, k0, k1):
    """-> (mask fp32 [n], thresh fp32 [2], kept int32 [n_layers])"""
    return load().thresh_mask(score, layer_off, layer_group, min_keep, int(k0), int(k1))


# ---- K3 ----------------------------------------------------------------------------------------
def channel_gather(src, out_idx=None, in_idx=None):
    return load().channel_gather(src, out_idx, in_idx)


def channel_gather_grouped(srcs, out_idx, in_idx):
    """out_idx / in_idx: lists with None for "keep all"."""
    dummy = None
    oi, ii, ho, hi = [], [], [], []
    for s, o, i in zip(srcs, out_idx, in_idx):
        if dummy is None:
            dummy = torch.empty(0, dtype=torch.int32, device=s.device)
        oi.append(dummy if o is None else o)
        ii.append(dummy if i is None else i)
        ho.append(0 if o is None else 1)
        hi.append(0 if i is None else 1)
    return load().channel_gather_grouped(list(srcs), oi, ii, ho, hi)


def bias_comp(W, act):
    return load().bias_comp(W, act)
