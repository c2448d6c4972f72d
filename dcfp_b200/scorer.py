"""Calibration scorer: BN hooks -> K1 class statistics -> (NCCL all-reduce) -> K2 EIC scores.

This is the GPU side of what `train.py --prune-type dcfp` produces in the reference
(train.py:215-216,255-270,286-287): the per-BN-channel importance `eic`.  The feature-map sized
work behind it -- `d(loss)/d(gamma_c) = sum_{n,h,w} dy * xhat` inside autograd's BN backward -- is
computed here by the label-keyed segmented reduction K1, which additionally resolves the sum by the
class of each pixel (north_star "class-conditional calibration statistics"):

    S1[k, c] = sum_{p: label(p) = k} v(p, c)        S2[k, c] = sum v^2        cnt[r, k] = #pixels

  mode "bwd":  v = dy * xhat   -> sum_k S1[k, c] == bn.weight.grad  (reference-exact EIC feed)
  mode "fwd":  v = BN output y -> class-conditional mean / variance of the feature map

Pixels whose label is outside [0, K) -- the ignore label 255 -- carry no loss but DO carry gradient at every layer
below the logits (receptive fields overlap them), and autograd's bn.weight.grad sums over them.  They are therefore
not discarded: the arenas hold K + 1 rows, row K collecting those pixels, so that the sum over all rows is the
reference's gradient exactly while rows [0, K) stay the class-conditional statistics (K <= 254).

HBM layout: all layers write into ONE fp64 *step arena* [2, K+1, sumC] (S1 rows,
then S2 rows; layer l owns columns [off_l, off_l + C_l)), so that
  * the end-of-step fold (dgamma = sum_k S1, totals += step, step = 0) is one launch,
  * the multi-GPU combine is one all-reduce of the *total arena* [2, K+1, sumC] + cnt [R, K].
Backward-mode launches are DEFERRED: the hook only records (x, dy, mean, invstd) -- tensors autograd
holds anyway, or that live for microseconds otherwise -- and `flush()` reduces every pending layer
in ONE grouped K1 launch once `flush_bytes` of feature maps are pending (180 GB of HBM make holding a
few GB of gradients free; a grouped launch runs at ~93 % of the HBM roofline, a 16 MB single-layer
launch at ~25-45 %).
"""
import os

import torch
import torch.nn as nn

from . import ops
from .lazy import PendingBN, deposit_shortcut_grad, take_shortcut_grad

MAX_RESOLUTIONS = 16


class _FusedBN(torch.autograd.Function):
    """Training-mode BatchNorm2d (+ the in-place ReLU that follows it) on the fused sm_100a kernels (SURVEY 8 f1,
    csrc/bn_fused.cu).  Replaces nn.BatchNorm2d -> nn.ReLU(inplace=True) of the reference nets
    (networks/backbone/resnet.py:26-56, networks/tools/aspp.py:15-24) for the duration of a scoring pass.  The
    backward reads (x, dy) ONCE for everything that is a sum -- the class rows S1/S2 of the scorer's arena, dgamma
    (their row sum: the reference's bn.weight.grad, pruners/dcfp_pruner.py:18) and dbeta -- and once more for dx.

    residual: the shortcut of a bottleneck block (resnet.py:49-56): y = relu(bn(x) + residual) in the same element-wise
    pass (lazy.PendingBN decides when).  The gate of that ReLU is y > 0, so the backward first forms dz = (y > 0) ? dy : 0
    (one torch kernel, what autograd's ReLU node did before) -- the shortcut's gradient -- and runs the un-gated passes on it."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, relu, residual=None, res_node=None):
        sc, module = layer.scorer, layer.module
        factor = 0.0
        rm = rv = None
        if module.track_running_stats and module.running_mean is not None:
            rm, rv = module.running_mean, module.running_var
            if module.momentum is None:  # cumulative moving average
                module.num_batches_tracked.add_(1)
                factor = 1.0 / float(module.num_batches_tracked)
            else:
                factor = module.momentum
                if sc.track_counters:
                    module.num_batches_tracked.add_(1)
        sums_f, sums_b = sc._bn_scratch(layer)
        t = sc._t_begin()
        y, mean, invstd = ops.bn_forward(x, weight, bias, rm, rv, sums_f, factor, module.eps, relu, workspace=sc._bn_workspace,
                                         residual=residual)
        sc._t_end(t, "bn_fwd", (2 if residual is None else 3) * x.numel() * x.element_size())
        if residual is None:
            ctx.save_for_backward(x, weight, bias, mean, invstd)
        else:
            assert relu, "a residual is only fused together with the ReLU that follows the sum"
            sc.fused_tail_calls += 1
            ctx.save_for_backward(x, weight, bias, mean, invstd, y)
        ctx.layer, ctx.relu, ctx.sums_b, ctx.has_res = layer, relu, sums_b, residual is not None
        # res_node: the backward node of the fused tail that PRODUCED the residual (the previous bottleneck): this block's
        # shortcut gradient is deposited there instead of going through autograd's accumulation (lazy.deposit_shortcut_grad)
        ctx.res_node = res_node if (residual is not None and ctx.needs_input_grad[5]) else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, weight, bias, mean, invstd = ctx.saved_tensors[:5]
        layer = ctx.layer
        sc = layer.scorer
        dy = dy.contiguous(memory_format=torch.channels_last)
        relu = ctx.relu
        deposit = take_shortcut_grad(ctx, sc._deposits)  # the next block's shortcut gradient, if it was handed over directly
        if ctx.has_res:  # gate on the stored output; dz is the gradient of the sum: the shortcut's gradient and BN's dy
            t = sc._t_begin()
            dy = ops.relu_grad(ctx.saved_tensors[5], dy, deposit)
            sc._t_end(t, "tail_relu_bwd", (3 if deposit is None else 4) * x.numel() * x.element_size())
            relu = False
        elif deposit is not None:
            dy = dy + deposit
        need_dx = ctx.needs_input_grad[0]
        keys = sc._keys_for(x.shape[2], x.shape[3])
        nb = x.numel() * x.element_size()
        if sc.timing:  # the two passes timed apart: B1 (class-keyed reduction) is the path's dominant kernel
            # an EMPTY event pair first: what two consecutive event records cost on this stream (inside a graph: one node
            # transition), so that a reader -- bench.py's roofline -- can take it out of the bracketed kernel times
            sc._t_end(sc._t_begin(), "event_pair", 0)
            t = sc._t_begin()
            ops.bn_backward(x, dy, weight, bias, mean, invstd, keys, layer.S1, layer.S2, sc.rows, ctx.sums_b, relu, need_dx, phases=1)
            sc._t_end(t, "bn_bwd_reduce", 2 * nb + keys.numel())
            t = sc._t_begin()
            dx, dgamma, dbeta = ops.bn_backward(x, dy, weight, bias, mean, invstd, keys, layer.S1, layer.S2, sc.rows, ctx.sums_b,
                                                relu, need_dx, phases=2)
            sc._t_end(t, "bn_bwd_dx", 3 * nb if need_dx else 0)
        else:
            dx, dgamma, dbeta = ops.bn_backward(x, dy, weight, bias, mean, invstd, keys, layer.S1, layer.S2, sc.rows, ctx.sums_b,
                                                relu, need_dx)
        sc.k1_bytes += 2 * nb + keys.numel()
        gres = None
        if ctx.has_res and ctx.needs_input_grad[5]:
            if ctx.res_node is not None:
                deposit_shortcut_grad(ctx.res_node, dy, sc._deposits)
            else:
                gres = dy
        return (dx if need_dx else None), dgamma, dbeta, None, None, gres, None


class _FusedLayer:
    __slots__ = ("scorer", "name", "module", "S1", "S2", "index")

    def __init__(self, scorer, name, module, S1, S2, index):
        self.scorer, self.name, self.module, self.S1, self.S2, self.index = scorer, name, module, S1, S2, index


def scored_layers(model):
    """BN layers the reference scores (dcfp_pruner.py:11-13), in `named_modules` order."""
    ignore = getattr(model, "ignore_prune_layer", [])
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)) and n not in ignore]


def shard_plan(n_images, micro_batch, world, rank):
    """Image indices rank `rank` scores at each step: micro-batches are fixed by GLOBAL image index
    ([0,mb), [mb,2mb), ...) and dealt round-robin, so the set of micro-batches -- and therefore every
    BN batch statistic -- does not depend on the number of GPUs.  Returns a list over steps of
    (lo, hi) slices; a trailing remainder that cannot fill `world` micro-batches is dropped."""
    if micro_batch < 1 or world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard plan arguments")
    n_steps = n_images // (micro_batch * world)
    return [((s * world + rank) * micro_batch, (s * world + rank + 1) * micro_batch) for s in range(n_steps)]


def average_over_ranks(t, group=None):
    """In-place mean over ranks of a small per-step vector (the reference's DDP averages gradients before
    `dcfp_pruning.step` sees them, engine.py:66).  One all-reduce; no-op without torch.distributed."""
    dist = torch.distributed
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(t, group=group)
            t /= world
    return t


class ClassStatsScorer:
    def __init__(self, model, num_classes, mode="bwd", r=0.999, process_group=None, flush_bytes=1 << 30, keep_totals=True,
                 timing=False, fused=True, track_counters=True, fuse_residual=True, shadow_unscored=True):
        ops.require_gpu()
        assert mode in ("bwd", "fwd")
        self.model, self.K, self.mode, self.r = model, int(num_classes), mode, r
        self.group = process_group
        self.layers = scored_layers(model)
        if not self.layers:
            raise ValueError("model has no scored BatchNorm layers")
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1 and \
                any(isinstance(m, nn.SyncBatchNorm) for _, m in self.layers):
            # engine.py:65 converts to SyncBN: those layers normalise with statistics of the GLOBAL batch, which neither
            # the hook path (rank-local mean / invstd) nor the fused kernels reproduce.  Refuse instead of scoring with
            # the wrong xhat (micro-batches are fixed by global index instead).
            raise RuntimeError("ClassStatsScorer: the model contains SyncBatchNorm layers and torch.distributed runs with world "
                               "size > 1; score a plain BatchNorm2d replica per rank (nn.SyncBatchNorm is not supported)")
        # fused: BN(+ReLU) forward / backward on the library's own kernels, the class-keyed sums coming out of the BN
        # backward itself (bwd mode, channels_last training-mode layers; everything else keeps torch's BN + the hooks)
        self.fused = bool(fused) and mode == "bwd"
        self.track_counters = bool(track_counters)
        self.device = self.layers[0][1].weight.device
        if self.device.type != "cuda":
            raise RuntimeError("ClassStatsScorer: the model must live on a CUDA device (no CPU fallback)")
        self.names = [n for n, _ in self.layers]
        sizes = [m.weight.numel() for _, m in self.layers]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.total_channels = self.offsets[-1]
        K, C = self.K, self.total_channels
        if not 1 <= K <= 254:
            raise ValueError("num_classes must be in [1, 254] (uint8 class keys, one more row for pixels outside [0, K))")
        R = self.rows = K + 1  # row K: pixels whose label is outside [0, K) (ignore label)
        self.step_arena = torch.zeros(2, R, C, dtype=torch.float64, device=self.device)
        # fused BN layers write their class rows into an fp32 twin of the step arena with 128-bit vector reductions (a
        # quarter of the atomic instructions); fp32 only across the CTAs of one launch: fold_step adds it to the fp64 totals
        # and zeroes it every step.  Needs 16-byte aligned rows: sum C and every layer's column offset multiples of 4.
        sizes_ = [m.weight.numel() for _, m in self.layers]
        self.step_arena32 = torch.zeros(2, R, C, dtype=torch.float32, device=self.device) \
            if (fused and mode == "bwd" and all(c % 4 == 0 for c in sizes_)) else None
        # pass-wide totals and pixel counts share one buffer: ONE all-reduce combines everything
        self.total_arena = torch.zeros(2 * R * C + MAX_RESOLUTIONS * K, dtype=torch.float64, device=self.device) if keep_totals else None
        self.totals = self.total_arena[:2 * R * C].view(2, R, C) if keep_totals else None
        self.cnt = (self.total_arena[2 * R * C:] if keep_totals else
                    torch.zeros(MAX_RESOLUTIONS * K, dtype=torch.float64, device=self.device)).view(MAX_RESOLUTIONS, K)
        self._views = {n: (self.step_arena[0][:, a:b], self.step_arena[1][:, a:b])
                       for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}
        self.resolutions = []  # (h, w) in discovery order -> row of `cnt`
        self._keys = {}
        self._labels = None
        self._handles = []
        self.flush_bytes = int(flush_bytes)
        self._pending = []
        self._pending_fwd = []
        self._pending_bytes = 0
        self.eic = torch.zeros(C, dtype=torch.float32, device=self.device)
        self._gamma = None
        self.steps = 0
        self.timing = timing
        self.k1_events = []  # (start, end, algorithmic bytes) per K1 launch when timing
        self.k1_bytes = 0
        self.phase_events = []  # (kind, start, end, algorithmic bytes) of the fused BN passes when timing
        self._phase_acc, self._k1_acc = {}, (0.0, 0, 0)  # totals over graph replays (accumulate_timing)
        # fused BN: one scratch per layer and direction (striped fp64 partial sums + coefficient vectors, csrc/bn_common.cuh),
        # all of them in one flat buffer zeroed once per step
        self._bn_ws = self._bn_workspace = None
        # BatchNorm2d layers the reference does NOT score (ignore_prune_layer: the last bottleneck's bn3, the ASPP / deep-supervision
        # heads) run on the same kernels -- they are part of the same forward / backward pass over the same kind of maps -- with
        # their class rows going to a scratch arena nobody reads
        scored = {id(m) for _, m in self.layers}
        self._shadow = [(n, m) for n, m in model.named_modules()
                        if type(m) is nn.BatchNorm2d and id(m) not in scored and m.weight is not None and m.weight.numel() % 4 == 0
                        and m.weight.device == self.device] if (self.fused and shadow_unscored) else []
        self._dump = None
        if self.fused:
            self._bn_ws_off = [0]
            for c in sizes + [m.weight.numel() for _, m in self._shadow]:
                self._bn_ws_off.append(self._bn_ws_off[-1] + ops.bn_scratch_elems(c))
            self._bn_ws = torch.zeros(2 * self._bn_ws_off[-1], dtype=torch.float64, device=self.device)
            # DCFP_BN_COOP=1: the one-launch cooperative forward (bit-reproducible statistics, no atomics; csrc/bn_coop.cuh)
            # instead of the statistics + normalise kernel pair, which measures faster on every c2 layer shape
            # (13.5 vs 23 us on a 17 MB layer, 36 vs 40 us on a 67 MB one).  Its per-CTA partial sums live in this
            # unzeroed workspace shared by all layers (stream-ordered).
            if os.environ.get("DCFP_BN_COOP", "0") == "1":
                self._bn_workspace = ops.bn_workspace(max(sizes), self.device)
        self._views32 = {n: (self.step_arena32[0][:, a:b], self.step_arena32[1][:, a:b])
                         for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])} if self.step_arena32 is not None else {}
        rows_of = self._views32 if self.step_arena32 is not None else self._views
        self._fused_layers = {n: _FusedLayer(self, n, m, rows_of[n][0], rows_of[n][1], i)
                              for i, (n, m) in enumerate(self.layers)} if self.fused else {}
        if self._shadow:
            cmax = max(m.weight.numel() for _, m in self._shadow)
            self._dump = torch.zeros(2, R, cmax, dtype=torch.float32 if self.step_arena32 is not None else torch.float64, device=self.device)
            for i, (n, m) in enumerate(self._shadow):
                c = m.weight.numel()
                self._fused_layers[n] = _FusedLayer(self, n, m, self._dump[0][:, :c], self._dump[1][:, :c], len(self.layers) + i)
        self._fused_calls = {}
        self.fused_tail_calls = 0  # fused BN calls that also added a shortcut and applied the ReLU behind it
        self._relu_after = {}  # bn name -> True once an in-place nn.ReLU was seen consuming that BN's output
        # bn name -> True once a ReLU was seen consuming (that BN's output + something): the BN then returns a lazy.PendingBN
        self._add_relu_after = {}
        self._deposits = [0]  # open shortcut-gradient deposits (lazy.deposit_shortcut_grad); must be 0 after a backward pass
        self.fuse_residual = bool(fuse_residual) and os.environ.get("DCFP_BN_FUSE_RESIDUAL", "1") != "0"
        self._patched = []
        self.fused_layer_calls = 0

    # ------------------------------------------------------------------ hooks
    def attach(self):
        for name, module in self.layers:
            self._handles.append(module.register_forward_hook(self._make_hook(name)))
        if self.fused:
            self._patch_modules()
        return self

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        self._pending, self._pending_fwd, self._pending_bytes = [], [], 0
        for module in self._patched:  # drop the instance-level forward: the class's own forward is back
            module.__dict__.pop("forward", None)
        self._patched = []

    # ------------------------------------------------------------------ fused BN (+ReLU)
    def _patch_modules(self):
        """Instance-level `forward` overrides, removed by detach():
          * every scored BatchNorm2d runs _FusedBN when its input is eligible (training mode, labels set, channels_last
            fp32 / bf16, >= 64 pixels), else its own forward (then the forward hook / K1 path scores it);
          * every nn.ReLU returns its input untouched when that input is a fused BN output that already went through
            the ReLU, and otherwise LEARNS: an in-place ReLU applied to a fused BN's output marks that BN as
            "followed by ReLU", so from the next step on the two are one kernel; a ReLU applied to (a fused BN's
            output + another tensor) -- the tail of a bottleneck, resnet.py:49-56 -- marks that BN as "feeds a residual
            sum", and from the next step on it returns a lazy.PendingBN, which becomes y = relu(bn(x) + shortcut) in one
            element-wise pass when the ReLU arrives (and the plain value if anything else does).
            The first step of a pass therefore runs BN, add and ReLU apart -- with bit-identical results (same fma,
            same rounding before the add, same gate)."""
        sc = self
        for name, module in self.layers + self._shadow:
            if "forward" in module.__dict__ or not isinstance(module, nn.BatchNorm2d) or module.weight is None:
                continue
            layer = self._fused_layers[name]

            def bn_forward(x, _m=module, _layer=layer, _name=name):
                if not sc._eligible(_m, x):
                    return type(_m).forward(_m, x)
                sc.fused_layer_calls += 1
                if sc._add_relu_after.get(_name, False):
                    def run(residual, relu, _x=x):
                        node = getattr(residual, "grad_fn", None)
                        if not (isinstance(node, _FusedBN._backward_cls) and node.has_res and node.layer.scorer is sc):
                            node = None  # the shortcut was not produced by a fused tail: its gradient goes through autograd
                        y = _FusedBN.apply(_x, _m.weight, _m.bias, _layer, relu, residual, node)
                        y._dcfp_bn = (_name, relu)
                        return y

                    return PendingBN.make(x, run, True, (_name, False))
                relu = sc._relu_after.get(_name, False)
                y = _FusedBN.apply(x, _m.weight, _m.bias, _layer, relu)
                y._dcfp_bn = (_name, relu)
                return y

            module.forward = bn_forward
            self._patched.append(module)
        for module in self.model.modules():
            if isinstance(module, nn.ReLU) and "forward" not in module.__dict__:
                def relu_forward(inp, _m=module):
                    if isinstance(inp, PendingBN):
                        return inp.relu(_m.inplace)
                    tag = getattr(inp, "_dcfp_bn", None)
                    if tag is not None:
                        if tag[1]:
                            return inp  # the fused BN kernel already applied the ReLU
                        if _m.inplace:
                            sc._relu_after[tag[0]] = True
                    elif sc.fuse_residual:
                        sc._learn_residual_tail(inp)
                    return type(_m).forward(_m, inp)

                module.forward = relu_forward
                self._patched.append(module)

    def check_deposits(self):
        """After a backward pass: every shortcut gradient a fused tail handed to its producer must have been consumed."""
        if self._deposits[0] != 0:
            n, self._deposits[0] = self._deposits[0], 0
            raise RuntimeError("dcfp_b200: %d shortcut gradient(s) deposited on a fused BN tail were never consumed (that tail's "
                               "backward did not run); score this model with fuse_residual=False" % n)

    def _learn_residual_tail(self, inp):
        """`inp` is about to go through a ReLU: if autograd says it is (fused BN output, no ReLU) + (anything), remember
        that BN (first such operand).  Reading the graph costs nothing on the device and needs no wrapper tensors."""
        node = getattr(inp, "grad_fn", None)
        if node is None or type(node).__name__ != "AddBackward0" or getattr(node, "_saved_alpha", 1) != 1:
            return
        for fn, _ in node.next_functions:
            layer = getattr(fn, "layer", None)  # _FusedBN's backward node carries what forward() put on ctx
            if isinstance(layer, _FusedLayer) and layer.scorer is self and not fn.relu and not fn.has_res:
                self._add_relu_after[layer.name] = True
                return

    def _eligible(self, module, x):
        if self._labels is None or not torch.is_grad_enabled() or not (module.training or module.running_mean is None):
            return False
        if module.weight.dtype != torch.float32 or not (x.requires_grad or module.weight.requires_grad):
            return False
        return ops.bn_supported(x)

    def _bn_scratch(self, layer):
        """(forward, backward) scratch of this call: slices of the per-step workspace, or fresh buffers when the module
        runs more than once per step (shared modules)."""
        n = self._fused_calls.get(layer.name, 0)
        self._fused_calls[layer.name] = n + 1
        if n > 0:
            C = layer.module.weight.numel()
            return ops.bn_scratch(C, self.device), ops.bn_scratch(C, self.device)
        a, b, half = self._bn_ws_off[layer.index], self._bn_ws_off[layer.index + 1], self._bn_ws_off[-1]
        return self._bn_ws[a:b], self._bn_ws[half + a:half + b]

    @staticmethod
    def _event():
        """A timing event; inside a CUDA-graph capture an EXTERNAL one (an event-record node: its timestamp is rewritten by
        every replay, so kernels can be timed in situ without the host's launch rate in the picture)."""
        e = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
        e.record()
        return e

    def _t_begin(self):
        return self._event() if self.timing else None

    def _t_end(self, e0, kind, nbytes):
        if e0 is None:
            return
        self.phase_events.append((kind, e0, self._event(), nbytes))

    def accumulate_timing(self):
        """After a synchronised graph replay: add the elapsed times of the captured (external) events to the running totals
        (`phase_times()` / `k1_time_ms()` then report the totals over all replays)."""
        for kind, e0, e1, nb in self.phase_events:
            ms, b, n = self._phase_acc.get(kind, (0.0, 0, 0))
            self._phase_acc[kind] = (ms + e0.elapsed_time(e1), b + nb, n + 1)
        ms, b, n = self._k1_acc
        self._k1_acc = (ms + sum(e0.elapsed_time(e1) for e0, e1, _ in self.k1_events), b + sum(x for _, _, x in self.k1_events),
                        n + len(self.k1_events))

    def reset_timing(self, drop_events=True):
        self._phase_acc, self._k1_acc = {}, (0.0, 0, 0)
        if drop_events:
            self.k1_events.clear()
            self.phase_events.clear()

    def set_labels(self, labels):
        """labels of the micro-batch about to run: [N, H0, W0] uint8 / int32 / int64 on the device."""
        self._labels = labels.contiguous()
        self._keys = {}
        if self._bn_ws is not None:
            self._bn_ws.zero_()
            self._fused_calls = {}
            if self._dump is not None:
                self._dump.zero_()

    def _keys_for(self, h, w):
        key = (h, w)
        if key not in self._keys:
            if key not in self.resolutions:
                if len(self.resolutions) >= MAX_RESOLUTIONS:
                    raise RuntimeError("more than %d distinct feature-map resolutions" % MAX_RESOLUTIONS)
                self.resolutions.append(key)
            row = self.cnt[self.resolutions.index(key)]
            self._keys[key] = ops.label_keys(self._labels, h, w, self.K, row)
        return self._keys[key]

    @staticmethod
    def _dense(t):
        return t if t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last) else t.contiguous()

    def _launch(self, items):
        """ONE K1 launch over `items` = [(x, dy|None, scale|None, shift|None, keys, S1, S2)], optionally timed."""
        nbytes = sum(x.numel() * x.element_size() * (2 if dy is not None else 1) + k.numel() for x, dy, _, _, k, _, _ in items)
        if self.timing:
            e0 = self._event()
        bwd = items[0][1] is not None
        affine = items[0][2] is not None
        # K + 1 classes for the kernel: label_keys maps every label outside [0, K) to the key K, which is row K here
        ops.class_stats_grouped([i[0] for i in items], [i[4] for i in items], self.rows, [i[5] for i in items], [i[6] for i in items],
                                dys=[i[1] for i in items] if bwd else None,
                                scales=[i[2] for i in items] if affine else None, shifts=[i[3] for i in items] if affine else None,
                                affine_mode=ops.AFFINE_INVSTD_MEAN if bwd else ops.AFFINE_SCALE_SHIFT)
        if self.timing:
            self.k1_events.append((e0, self._event(), nbytes))
        self.k1_bytes += nbytes

    def _flush_fwd(self):
        """Deferred forward functor: y = (x - mean) * invstd * gamma + beta = x * scale + shift with the per-channel
        scale / shift of ALL pending layers computed by four small launches on the concatenated vectors."""
        items = self._pending_fwd
        self._pending_fwd = []
        with torch.no_grad():
            mean = torch.cat([i[1] for i in items])
            invstd = torch.cat([i[2] for i in items])
            gamma = torch.cat([i[3].detach().float() for i in items])
            beta = torch.cat([i[4].detach().float() for i in items])
            scale = invstd * gamma
            shift = beta - mean * scale
        launch, pos = [], 0
        for xd, m, _, _, _, keys, S1, S2 in items:
            c = m.numel()
            launch.append((xd, None, scale[pos:pos + c], shift[pos:pos + c], keys, S1, S2))
            pos += c
        groups = {}
        for it in launch:
            groups.setdefault((it[0].dtype, it[0].is_contiguous()), []).append(it)
        for its in groups.values():
            self._launch(its)

    def flush(self):
        """Reduce every pending (deferred) layer; grouped by (dtype, layout) as one launch needs."""
        if self._pending_fwd:
            self._pending_bytes = 0
            self._flush_fwd()
        if not self._pending:
            return
        groups = {}
        for item in self._pending:
            x = item[0]
            groups.setdefault((x.dtype, x.is_contiguous()), []).append(item)
        self._pending, self._pending_bytes = [], 0
        for items in groups.values():
            self._launch(items)

    def _make_hook(self, name):
        S1, S2 = self._views[name]

        def hook(module, inputs, output):
            if self._labels is None or getattr(output, "_dcfp_bn", None) is not None:  # fused layers score themselves
                return
            x = inputs[0]
            if self.mode == "fwd":
                # v = y (the BN output, pre-ReLU).  The in-place ReLU that follows overwrites y, so either reduce it
                # right here (one launch per layer), or -- when autograd recorded the batch statistics -- DEFER: keep
                # (x, mean, invstd) and let ONE grouped launch at the end of the forward pass evaluate
                # y = (x - mean) * invstd * gamma + beta inside K1's value functor.
                node = output.grad_fn if (self.flush_bytes > 0 and torch.is_grad_enabled()) else None
                mean = getattr(node, "_saved_result1", None) if node is not None else None
                invstd = getattr(node, "_saved_result2", None) if node is not None else None
                del node
                training = module.training or module.running_mean is None
                if self.flush_bytes > 0 and not training:
                    mean, invstd = module.running_mean.float(), torch.rsqrt(module.running_var.float() + module.eps)
                if mean is not None and invstd is not None and mean.numel() == x.shape[1] and mean.dtype == torch.float32:
                    xd = self._dense(x.detach())
                    self._pending_fwd.append((xd, mean, invstd, module.weight, module.bias,
                                              self._keys_for(xd.shape[2], xd.shape[3]), S1, S2))
                    self._pending_bytes += xd.numel() * xd.element_size()
                    if self._pending_bytes >= self.flush_bytes:
                        self.flush()
                    return
                y = self._dense(output.detach())
                self._launch([(y, None, None, None, self._keys_for(y.shape[2], y.shape[3]), S1, S2)])
                return
            if not output.requires_grad:
                return
            training = module.training or module.running_mean is None
            # batch statistics autograd's BN node saved (CudnnBatchNormBackward0 / NativeBatchNormBackward0: result1 =
            # mean, result2 = invstd).  Read them NOW: capturing the node in the gradient hook below would close a
            # reference cycle node -> hook -> closure -> node that Python's GC cannot see, leaking x every step.
            mean = invstd = None
            if training:
                node = output.grad_fn
                mean = getattr(node, "_saved_result1", None)
                invstd = getattr(node, "_saved_result2", None)
                del node
                if mean is None or invstd is None or mean.numel() != x.shape[1] or mean.dtype != torch.float32:
                    mean = invstd = None

            def on_grad(dy, mean=mean, invstd=invstd):
                with torch.no_grad():
                    xd = x.detach()
                    if not training:
                        mean = module.running_mean.float()
                        invstd = torch.rsqrt(module.running_var.float() + module.eps)
                    elif mean is None:
                        # BN implementation that saves nothing usable: recompute the batch statistics
                        var, mean = torch.var_mean(xd.float(), dim=(0, 2, 3), unbiased=False)
                        invstd = torch.rsqrt(var + module.eps)
                    # K1 wants x and dy in ONE layout: follow x (the tensor autograd saved, never copied)
                    if xd.is_contiguous():
                        g = dy.contiguous()
                    elif xd.is_contiguous(memory_format=torch.channels_last):
                        g = dy.contiguous(memory_format=torch.channels_last)
                    else:
                        xd, g = xd.contiguous(), dy.contiguous()
                    item = (xd, g, invstd.contiguous(), mean.contiguous(), self._keys_for(xd.shape[2], xd.shape[3]), S1, S2)
                    if self.flush_bytes <= 0:
                        self._launch([item])
                        return
                    self._pending.append(item)
                    self._pending_bytes += 2 * xd.numel() * xd.element_size()
                    if self._pending_bytes >= self.flush_bytes:
                        self.flush()

            output.register_hook(on_grad)

        return hook

    # ------------------------------------------------------------------ reductions
    def fold_step(self):
        """End of one step: flush deferred layers, then ONE launch yields dgamma = sum_k S1 (fp32 [sumC]),
        adds the step arena into the pass totals and zeroes it for the next step."""
        self.check_deposits()
        self.flush()
        return ops.fold_step(self.step_arena, self.totals, self.step_arena32)

    def step_rows(self, name):
        """(S1, S2) fp64 [K+1, C] of layer `name` accumulated since the last fold (hook-fed fp64 rows + fused fp32 rows)."""
        a, b = self._views[name]
        if name in self._views32:
            a, b = a + self._views32[name][0].double(), b + self._views32[name][1].double()
        return a, b

    def all_reduce_totals(self):
        """ONE collective for all layers, both moments and the pixel counts (SUM, fp64, NCCL over NVLink)."""
        dist = torch.distributed
        if self.total_arena is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.total_arena, group=self.group)

    def gamma(self):
        if self._gamma is None:
            self._gamma = torch.cat([m.weight.detach().reshape(-1) for _, m in self.layers]).float()
        return self._gamma

    def eic_step(self, dgamma):
        """One EIC update (dcfp_pruner.py:15-20) from the (rank-averaged) BN-gamma gradient of this step."""
        ops.eic_update_flat(dgamma, self.gamma(), self.eic, self.r, first_step=(self.steps == 0))
        self.steps += 1

    def eic_dict(self):
        """`{'eic': {bn_name: Tensor[C]}}` -- the layout of score.pth (dcfp_pruner.py:10,25-26)."""
        return {"eic": {n: self.eic[a:b].clone() for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}}

    def class_stats(self):
        """Pass totals {name: (S1[K,C], S2[K,C])} plus per-resolution pixel counts {(h,w): cnt[K]}; pixels outside
        [0, K) (ignore label) are not in any class: their sums are `outside_stats()`."""
        src = self.totals if self.totals is not None else self.step_arena
        K = self.K
        views = {n: (src[0][:K, a:b], src[1][:K, a:b]) for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}
        return views, {r: self.cnt[i] for i, r in enumerate(self.resolutions)}

    def outside_stats(self):
        """{name: (S1[C], S2[C])} of the pixels whose label is outside [0, K): class sums + these = sums over all pixels."""
        src = self.totals if self.totals is not None else self.step_arena
        K = self.K
        return {n: (src[0][K, a:b], src[1][K, a:b]) for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}

    def k1_time_ms(self):
        """(sum of K1 launch durations in ms, algorithmic bytes, launches) -- call after a synchronize."""
        if self._k1_acc[2]:
            return self._k1_acc
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in self.k1_events)
        return ms, sum(b for _, _, b in self.k1_events), len(self.k1_events)

    def phase_times(self):
        """{kind: (ms, algorithmic bytes, calls)} of the fused BN passes (timing=True) -- call after a synchronize.
        kinds: bn_fwd (statistics + normalise), bn_bwd_reduce (B1: the class-keyed reduction), bn_bwd_dx (B2), tail_relu_bwd
        (ReLU backward + gradient fan-in of a bottleneck tail), event_pair
        (an empty bracket recorded in front of every B1 bracket: the cost of the timing events themselves)."""
        if self._phase_acc:
            return dict(self._phase_acc)
        out = {}
        for kind, e0, e1, nb in self.phase_events:
            ms, b, n = out.get(kind, (0.0, 0, 0))
            out[kind] = (ms + e0.elapsed_time(e1), b + nb, n + 1)
        return out


class CalibrationRun:
    """Step-wise driver of the scoring pass (bench.py times `step`; `score_calibration_set` loops it).

    Protocol (oracle/scoring_ref.py follows the same one): for every micro-batch
    (fixed by global image index)  zero_grad -> loss = model(x, y, deepsup=True) -> backward -> EIC step
    on this step's gradient; BN runs in train mode on the micro-batch like the reference's training
    step (train.py:255-268), no optimizer step, running statistics restored at `close`.  With
    torch.distributed initialised every step's dgamma vector is averaged over the ranks with one
    all-reduce before the sign gate (the reference gates on the DDP-averaged gradient, engine.py:66)."""

    def __init__(self, model, num_classes, r=0.999, mode="bwd", restore_bn_stats=True, flush_bytes=1 << 30, keep_totals=True,
                 timing=False, process_group=None, seed=None, scores_only=False, fused=True, graph=True, autocast_dtype=None,
                 fuse_residual=True):
        ops.require_gpu()
        self.model = model
        self.seed = seed
        self.scores_only = bool(scores_only)
        self.restore_bn_stats = bool(restore_bn_stats)
        # autocast_dtype=torch.bfloat16: the forward runs under torch.autocast -- bf16 convolutions and bf16 feature maps, which
        # the fused BN kernels / K1 read and write as bf16 (fp32 statistics and sums).  Not the reference's arithmetic (fp32).
        self.autocast_dtype = autocast_dtype
        self._frozen = []
        self.device = next(model.parameters()).device
        # restore_bn_stats puts num_batches_tracked back at close(): the fused layers then skip its per-layer increment
        self.scorer = ClassStatsScorer(model, num_classes, mode=mode, r=r, process_group=process_group, flush_bytes=flush_bytes,
                                       keep_totals=keep_totals, timing=timing, fused=fused,
                                       track_counters=not restore_bn_stats, fuse_residual=fuse_residual)
        self._saved = None
        self.group = process_group
        self.closed = True
        # CUDA graph of one step (forward + backward + fold): a c2 step is ~900 launches issued from Python and the autograd
        # thread; replaying them from a graph takes the host out of the critical path.  Captured on the third step of a
        # given input shape (the first learns the BN->ReLU pairs and lets cuDNN autotune, the second runs the final kernels
        # once eagerly); any failure to capture falls back to eager launches of the SAME kernels.
        self.graph = bool(graph) and mode == "bwd"
        self._graph = None       # (key, CUDAGraph, static x, static y, dgamma, loss)
        self._graph_seen = {}    # input signature -> eager steps run
        self._graph_grads = []   # (parameter, its .grad tensor inside the graph's pool)
        self.graph_replays = 0
        self.graph_launches = 0
        dist = torch.distributed
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self._comm_stream = torch.cuda.Stream(device=self.device) if (multi and mode == "bwd") else None
        self._open()

    def _open(self):
        """Puts the model into scoring state: (scores_only) freeze non-BN parameters, attach hooks / fused forwards, snapshot the
        BN running statistics, train mode."""
        model = self.model
        # scores_only: the EIC needs d(loss)/d(gamma) of the BN layers only, and K1 computes it from (x, dy) itself.
        # Freezing every non-BN parameter keeps the activation-gradient chain (dgrad) but drops the weight-gradient
        # convolutions (wgrad, ~17 % of a c2 step) that the reference's training loop needs for its optimizer and a
        # calibration pass does not.  Off by default: the default step does the reference's full backward.
        if self.scores_only:
            bn_params = {id(p) for m in model.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm) for p in m.parameters()}
            for p_ in model.parameters():
                if id(p_) not in bn_params and p_.requires_grad:
                    p_.requires_grad_(False)
                    self._frozen.append(p_)
        self.scorer.attach()
        if self.restore_bn_stats:  # train-mode BN updates its running statistics: snapshot them (a few fused launches)
            bns = [m for m in model.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.running_mean is not None]
            live = [m.running_mean for m in bns] + [m.running_var for m in bns]
            counters = [m.num_batches_tracked for m in bns]
            self._saved = (live, torch._foreach_mul(live, 1.0) if live else [], counters,
                           torch._foreach_add(counters, 0) if counters else [])
        self._was_training = model.training
        model.train()
        for p_, g in self._graph_grads:  # a kept graph writes the gradients into these tensors
            p_.grad = g
        self.closed = False

    def reopen(self):
        """A closed run whose CUDA graph was kept (close(keep_graph=True)) scores again from a clean state."""
        sc = self.scorer
        sc.steps = 0
        sc.eic.zero_()
        sc.step_arena.zero_()
        if sc.step_arena32 is not None:
            sc.step_arena32.zero_()
        if sc.total_arena is not None:
            sc.total_arena.zero_()
        else:
            sc.cnt.zero_()
        sc._pending, sc._pending_fwd, sc._pending_bytes = [], [], 0
        sc.k1_events.clear()
        sc.phase_events.clear()
        self._open()

    def _forward_backward(self, x, y):
        sc = self.scorer
        sc.set_labels(y)
        self.model.zero_grad(set_to_none=True)
        if self.autocast_dtype is not None:
            with torch.autocast(device_type="cuda", dtype=self.autocast_dtype):
                out = self.model(x, y if y.dtype == torch.long else y.long(), deepsup=True)
        else:
            out = self.model(x, y if y.dtype == torch.long else y.long(), deepsup=True)
        loss = out["loss"] if isinstance(out, dict) else out
        if sc.mode == "bwd":
            loss.backward()
            sc.check_deposits()
        return sc.fold_step(), loss.detach().float()

    def _capture(self, key, x, y):
        gx, gy = x.clone(memory_format=torch.preserve_format), y.clone()
        g = torch.cuda.CUDAGraph()
        if self.scorer.timing:  # only the events recorded inside the capture are re-stamped by the replays
            self.scorer.reset_timing()
        launches0 = ops.launch_count()
        try:
            with torch.cuda.graph(g):
                dgamma, loss = self._forward_backward(gx, gy)
        except Exception as e:  # not capturable (data-dependent control flow, an op that synchronises, ...): eager from now on
            import warnings
            warnings.warn("dcfp_b200: the scoring step could not be captured into a CUDA graph (%s: %s); launching eagerly" %
                          (type(e).__name__, str(e).split("\n")[0][:200]))
            self.graph = False
            self.scorer._pending, self.scorer._pending_fwd, self.scorer._pending_bytes = [], [], 0
            torch.cuda.synchronize()
            return False
        self._graph = (key, g, gx, gy, dgamma, loss)
        self._graph_grads = [(p_, p_.grad) for p_ in self.model.parameters() if p_.grad is not None]
        self.graph_launches = ops.launch_count() - launches0  # kernels of this library inside one replay
        return True

    def step(self, x, y, mb_index=None):
        """x [N,3,H,W] fp32, y [N,H,W] integer labels, both ON THE DEVICE.  Returns the loss tensor (device).
        mb_index: GLOBAL micro-batch index; with `seed` set the RNG (Dropout2d of the deep-supervision head,
        deeplabv3.py:40) is keyed by it so that results do not depend on the number of ranks."""
        sc = self.scorer
        if self.seed is not None and mb_index is not None:
            torch.manual_seed(self.seed + int(mb_index))
        dgamma = None
        if self.graph:
            key = (tuple(x.shape), x.dtype, tuple(x.stride()), tuple(y.shape), y.dtype, sc.timing)
            if self._graph is not None and self._graph[0] == key:
                _, g, gx, gy, dgamma, loss = self._graph
                gx.copy_(x, non_blocking=True)
                gy.copy_(y, non_blocking=True)
                g.replay()
                self.graph_replays += 1
                if sc.timing:  # in-graph (external) events: read them once the replay has finished
                    torch.cuda.synchronize()
                    sc.accumulate_timing()
            else:
                seen = self._graph_seen.get(key, 0)
                self._graph_seen[key] = seen + 1
                if seen >= 2 and self._capture(key, x, y):  # the captured launches have not run yet: replay them once
                    _, g, gx, gy, dgamma, loss = self._graph
                    g.replay()
                    self.graph_replays += 1
                    if sc.timing:
                        torch.cuda.synchronize()
                        sc.accumulate_timing()
        if dgamma is None:
            dgamma, loss = self._forward_backward(x, y)
        else:
            dgamma, loss = dgamma.clone(), loss.clone()  # the graph's outputs are overwritten by the next replay
        if sc.mode == "bwd":
            if self._comm_stream is None:
                average_over_ranks(dgamma, self.group)
                sc.eic_step(dgamma)
            else:
                # N > 1: the all-reduce of this step's dgamma and the EIC update it feeds run on a side stream -- the score
                # state only has to be final when it is read, so the next step's forward does not wait for the slowest rank
                # of this one (ranks may drift by up to a step instead of meeting at every all-reduce)
                main = torch.cuda.current_stream(self.device)
                self._comm_stream.wait_stream(main)
                with torch.cuda.stream(self._comm_stream):
                    average_over_ranks(dgamma, self.group)
                    sc.eic_step(dgamma)
                dgamma.record_stream(self._comm_stream)
        return loss

    def sync_scores(self):
        """Makes the current stream wait for the side-stream EIC updates (N > 1); call before reading `scorer.eic`."""
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def close(self, keep_graph=False):
        """Restores the model (hooks off, running statistics, train / eval mode, requires_grad).  keep_graph: the captured
        step and its memory pool survive for `reopen()` (score_calibration_set's run cache)."""
        if not keep_graph:
            self._graph, self._graph_grads = None, []  # releases the graph's private memory pool
        if self.closed:
            return
        self.sync_scores()
        self.closed = True
        for p_ in self._frozen:
            p_.requires_grad_(True)
        self._frozen = []
        self.scorer.detach()
        self.model.train(self._was_training)
        self.model.zero_grad(set_to_none=True)
        if self._saved is not None:
            live, snap, counters, csnap = self._saved
            with torch.no_grad():
                if live:
                    torch._foreach_copy_(live, snap)
                if counters:
                    torch._foreach_copy_(counters, csnap)
            self._saved = None


#: score_calibration_set keeps the last CalibrationRun of a model -- with its captured CUDA graph and the graph's memory pool
#: (~20 GB for a 512x1024 ResNet-101 micro-batch) -- so that the next call on the same model with the same settings skips
#: capture.  Keyed by the model's identity AND the addresses of its parameters / buffers: a model whose tensors were
#: re-allocated never sees a stale graph.  `release_cached_runs()` frees everything.
_RUN_CACHE = {}


def _run_key(model, *settings):
    ptrs = tuple(t.data_ptr() for t in list(model.parameters()) + list(model.buffers()))
    return settings + (hash(ptrs), len(ptrs))


def release_cached_runs(model=None):
    """Drops the cached CalibrationRun (CUDA graph + its memory pool) of `model`, or of every model."""
    for k in [k for k in _RUN_CACHE if model is None or k == id(model)]:
        ref, _, run = _RUN_CACHE.pop(k)
        run.close()


def release_cached_runs_by_id(mid):
    entry = _RUN_CACHE.pop(mid, None)
    if entry is not None:
        entry[2]._graph, entry[2]._graph_grads = None, []


def score_calibration_set(model, images, labels, num_classes, micro_batch=2, r=0.999, restore_bn_stats=True, flush_bytes=1 << 30,
                          return_class_stats=False, seed=0, scores_only=False, channels_last=True, fused=True, n_images=None, graph=True):
    """Public end-to-end call: HOST images [n,3,H,W] / labels [n,H,W] -> EIC scores on the host.

    `images` may also be a callable `fetch(lo, hi) -> (x [hi-lo,3,H,W], y [hi-lo,H,W])` returning host tensors for the global
    image range [lo, hi) -- a loader for calibration sets that do not fit in host memory at once (then `labels` is ignored
    and `n_images` gives the size of the set).

    Every step copies its micro-batch host->device from pinned memory and reads the step's loss back;
    with torch.distributed initialised the micro-batches are dealt round-robin (`shard_plan`).  Returns
    {'eic': {bn_name: FloatTensor[C] (cpu)}, '_stats': {...}} -- `{'eic': ...}` is score.pth's layout."""
    ops.require_gpu()
    device = next(model.parameters()).device
    # channels_last is cuDNN's native tensor-core layout (no per-convolution transposes: 52 -> 36 ms per c2 step) and
    # takes K1's NHWC path.  An NCHW model is converted for the duration of the pass (strides only: parameter VALUES do
    # not change) and converted back afterwards; channels_last=False scores it as it is.
    def is_cl(p):
        return p.dim() == 4 and p.shape[1] > 1 and p.shape[2] * p.shape[3] > 1 and not p.is_contiguous() \
            and p.is_contiguous(memory_format=torch.channels_last)
    nhwc = any(is_cl(p) for p in model.parameters())
    converted = False
    if channels_last and not nhwc:
        model.to(memory_format=torch.channels_last)
        nhwc = converted = True
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    world = torch.distributed.get_world_size() if dist_on else 1
    rank = torch.distributed.get_rank() if dist_on else 0
    fetch = images if callable(images) else None
    if fetch is not None:
        if n_images is None:
            raise ValueError("score_calibration_set: n_images is required when `images` is a loader")
        x0, y0 = fetch(0, micro_batch)
        img_shape, img_dtype, lab_shape, lab_dtype = tuple(x0.shape[1:]), x0.dtype, tuple(y0.shape[1:]), y0.dtype
    else:
        n_images = images.shape[0]
        img_shape, img_dtype, lab_shape, lab_dtype = tuple(images.shape[1:]), images.dtype, tuple(labels.shape[1:]), labels.dtype
    plan = shard_plan(n_images, micro_batch, world, rank)
    key = _run_key(model, int(num_classes), micro_batch, float(r), bool(restore_bn_stats), int(flush_bytes), bool(return_class_stats),
                   seed, bool(scores_only), bool(fused), world, img_shape, lab_shape, str(img_dtype), str(lab_dtype))
    cached = _RUN_CACHE.pop(id(model), None) if graph else None
    run = None
    if cached is not None:
        if cached[0]() is model and cached[1] == key and cached[2].closed:
            run = cached[2]
            run.reopen()
        else:
            cached[2].close()
    if run is None:
        run = CalibrationRun(model, num_classes, r=r, restore_bn_stats=restore_bn_stats, flush_bytes=flush_bytes,
                             keep_totals=return_class_stats, seed=seed, scores_only=scores_only, fused=fused, graph=graph)
    h2d = d2h = 0
    losses = torch.empty(max(len(plan), 1), dtype=torch.float32).pin_memory()
    launches0 = ops.launch_count()
    try:
        # host->device copies run one step ahead on their own stream into two fixed staging buffers (pinned source,
        # copy engine): step s+1's micro-batch is already resident when step s's backward finishes.  Fixed buffers,
        # fenced by events, keep the caching allocator out of it (per-step cross-stream allocations made it re-grow).
        copy_stream = torch.cuda.Stream(device=device)
        main_stream = torch.cuda.current_stream(device)
        stage = [(torch.empty((micro_batch,) + img_shape, dtype=img_dtype, device=device),
                  torch.empty((micro_batch,) + lab_shape, dtype=lab_dtype, device=device)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        pinned = [None, None]

        def upload(step):
            lo, hi = plan[step]
            b = step & 1
            xb, yb = fetch(lo, hi) if fetch is not None else (images[lo:hi], labels[lo:hi])
            if not xb.is_pinned():
                xb, yb = xb.pin_memory(), yb.pin_memory()
            pinned[b] = (xb, yb)  # keep the pinned source alive until the copy has run
            with torch.cuda.stream(copy_stream):
                if step >= 2:
                    copy_stream.wait_event(consumed[b])
                stage[b][0].copy_(xb, non_blocking=True)
                stage[b][1].copy_(yb, non_blocking=True)
                ready[b].record(copy_stream)
            return xb.numel() * xb.element_size() + yb.numel() * yb.element_size()

        pending_bytes = upload(0) if plan else 0
        for step, (lo, hi) in enumerate(plan):
            b = step & 1
            h2d += pending_bytes
            if step + 1 < len(plan):
                pending_bytes = upload(step + 1)
            main_stream.wait_event(ready[b])
            x, y = stage[b]
            if nhwc:
                x = x.contiguous(memory_format=torch.channels_last)
            loss = run.step(x, y, mb_index=lo // micro_batch)
            consumed[b].record(main_stream)
            losses[step:step + 1].copy_(loss.reshape(1), non_blocking=True)
            d2h += 4
        run.sync_scores()
        if return_class_stats:
            run.scorer.all_reduce_totals()
    finally:
        keep = bool(graph) and run._graph is not None and not converted
        run.close(keep_graph=keep)
        if keep:
            import weakref
            mid = id(model)
            _RUN_CACHE[mid] = (weakref.ref(model, lambda _r, mid=mid: release_cached_runs_by_id(mid)), key, run)
        if converted:
            model.to(memory_format=torch.contiguous_format)
    sc = run.scorer
    flat = sc.eic.cpu()  # ONE device->host copy; the per-layer tensors of score.pth are slices of it
    out = {"eic": {n: flat[a:b].clone() for n, a, b in zip(sc.names, sc.offsets[:-1], sc.offsets[1:])}}
    d2h += sc.eic.numel() * 4
    if return_class_stats:
        totals = (sc.totals if sc.totals is not None else sc.step_arena).cpu()
        cnt = sc.cnt.cpu()
        K = sc.K  # rows [0, K): the classes; row K: pixels whose label is outside [0, K) (ignore label)
        out["class_stats"] = {n: (totals[0][:K, a:b].clone(), totals[1][:K, a:b].clone())
                              for n, a, b in zip(sc.names, sc.offsets[:-1], sc.offsets[1:])}
        out["outside_stats"] = {n: (totals[0][K, a:b].clone(), totals[1][K, a:b].clone())
                                for n, a, b in zip(sc.names, sc.offsets[:-1], sc.offsets[1:])}
        out["class_counts"] = {r: cnt[i].clone() for i, r in enumerate(sc.resolutions)}
    torch.cuda.synchronize(device)
    out["_stats"] = dict(steps=len(plan), h2d_bytes=h2d, d2h_bytes=d2h, launches=ops.launch_count() - launches0,
                         losses=losses[:len(plan)].clone())
    return out
