"""Calibration scorer: BN hooks -> K1 class statistics -> (NCCL all-reduce) -> K2 EIC scores.

This is the GPU side of what `train.py --prune-type dcfp` produces in the reference
(train.py:215-216,255-270,286-287): the per-BN-channel importance `eic`.  The feature-map sized
work behind it -- `d(loss)/d(gamma_c) = sum_{n,h,w} dy * xhat` inside autograd's BN backward -- is
computed here by the label-keyed segmented reduction K1, which additionally resolves the sum by the
class of each pixel (north_star "class-conditional calibration statistics"):

    S1[k, c] = sum_{p: label(p) = k} v(p, c)        S2[k, c] = sum v^2        cnt[r, k] = #pixels

  mode "bwd":  v = dy * xhat   -> sum_k S1[k, c] == bn.weight.grad  (reference-exact EIC feed)
  mode "fwd":  v = BN output y -> class-conditional mean / variance of the feature map

All layers write into ONE fp64 arena [S1 (K x sumC) | S2 (K x sumC) | cnt (R x K)] so that the
multi-GPU combine is a single all-reduce and the class reduction / EIC update are one launch each.
"""
import torch
import torch.nn as nn

from . import ops

MAX_RESOLUTIONS = 16


def scored_layers(model):
    """BN layers the reference scores (dcfp_pruner.py:11-13), in `named_modules` order."""
    ignore = getattr(model, "ignore_prune_layer", [])
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)) and n not in ignore]


class ClassStatsScorer:
    def __init__(self, model, num_classes, mode="bwd", r=0.999, process_group=None):
        ops.require_gpu()
        assert mode in ("bwd", "fwd")
        self.model, self.K, self.mode, self.r = model, int(num_classes), mode, r
        self.group = process_group
        self.layers = scored_layers(model)
        if not self.layers:
            raise ValueError("model has no scored BatchNorm layers")
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
        self.arena = torch.zeros(2 * K * C + MAX_RESOLUTIONS * K, dtype=torch.float64, device=self.device)
        self.S1 = self.arena[:K * C].view(K, C)
        self.S2 = self.arena[K * C:2 * K * C].view(K, C)
        self.cnt = self.arena[2 * K * C:].view(MAX_RESOLUTIONS, K)
        self._views = {n: (self.S1[:, a:b], self.S2[:, a:b]) for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}
        self.resolutions = []  # (h, w) in discovery order -> row of `cnt`
        self._keys = {}
        self._labels = None
        self._handles = []
        self.eic = torch.zeros(C, dtype=torch.float32, device=self.device)
        self.steps = 0
        self.launches = 0

    # ------------------------------------------------------------------ hooks
    def attach(self):
        for name, module in self.layers:
            self._handles.append(module.register_forward_hook(self._make_hook(name)))
        return self

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def set_labels(self, labels):
        """labels of the micro-batch about to run: [N, H0, W0] uint8 / int32 / int64 on the device."""
        self._labels = labels.contiguous()
        self._keys = {}

    def _keys_for(self, h, w):
        key = (h, w)
        if key not in self._keys:
            if key not in self.resolutions:
                if len(self.resolutions) >= MAX_RESOLUTIONS:
                    raise RuntimeError("more than %d distinct feature-map resolutions" % MAX_RESOLUTIONS)
                self.resolutions.append(key)
            row = self.cnt[self.resolutions.index(key)]
            self._keys[key] = ops.label_keys(self._labels, h, w, self.K, row)
            self.launches += 1
        return self._keys[key]

    def _make_hook(self, name):
        S1, S2 = self._views[name]

        def hook(module, inputs, output):
            if self._labels is None:
                return
            x = inputs[0]
            if self.mode == "fwd":
                y = output.detach()
                y = y if y.is_contiguous() or y.is_contiguous(memory_format=torch.channels_last) else y.contiguous()
                ops.class_stats(y, self._keys_for(y.shape[2], y.shape[3]), self.K, S1, S2)
                self.launches += 1
                return
            if not output.requires_grad:
                return
            node = output.grad_fn
            training = module.training or module.running_mean is None

            def on_grad(dy):
                with torch.no_grad():
                    mean = getattr(node, "_saved_result1", None) if training else None
                    invstd = getattr(node, "_saved_result2", None) if training else None
                    xd = x.detach()
                    if not training:
                        mean = module.running_mean
                        invstd = torch.rsqrt(module.running_var + module.eps)
                    elif mean is None or invstd is None or mean.numel() != xd.shape[1]:
                        var, mean = torch.var_mean(xd.float(), dim=(0, 2, 3), unbiased=False)
                        invstd = torch.rsqrt(var + module.eps)
                    scale = invstd.float().contiguous()
                    shift = (-mean.float() * scale).contiguous()
                    g = dy if dy.is_contiguous() or dy.is_contiguous(memory_format=torch.channels_last) else dy.contiguous()
                    if xd.stride() != g.stride():
                        xd, g = xd.contiguous(), g.contiguous()
                    ops.class_stats(xd, self._keys_for(xd.shape[2], xd.shape[3]), self.K, S1, S2, dy=g, scale=scale, shift=shift)
                    self.launches += 1

            output.register_hook(on_grad)

        return hook

    # ------------------------------------------------------------------ reductions
    def zero_stats(self):
        self.arena.zero_()

    def all_reduce(self):
        """ONE collective for all layers, classes and counts (SUM, fp64, NCCL over NVLink)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            if torch.distributed.get_world_size(self.group) > 1:
                torch.distributed.all_reduce(self.arena, group=self.group)

    def dgamma(self):
        """sum_k S1[k, c] for every scored channel, fp32 [sumC] -- ONE launch over the shared arena."""
        self.launches += 1
        return ops.reduce_classes(self.S1)

    def eic_step(self, dgamma=None, gamma=None):
        """One EIC update (dcfp_pruner.py:15-20) from the class-resolved BN-gamma gradient."""
        dgamma = self.dgamma() if dgamma is None else dgamma
        if gamma is None:
            gamma = torch.cat([m.weight.detach().reshape(-1) for _, m in self.layers]).float()
        ops.eic_update_flat(dgamma, gamma, self.eic, self.r, first_step=(self.steps == 0))
        self.launches += 1
        self.steps += 1
        return dgamma

    def eic_dict(self):
        """`{'eic': {bn_name: Tensor[C]}}` -- the layout of score.pth (dcfp_pruner.py:10,25-26)."""
        return {"eic": {n: self.eic[a:b].clone() for n, a, b in zip(self.names, self.offsets[:-1], self.offsets[1:])}}

    def class_stats(self):
        """{name: (S1[K,C], S2[K,C])} views plus per-resolution counts."""
        return dict(self._views), {r: self.cnt[i] for i, r in enumerate(self.resolutions)}


def score_calibration_set(model, images, labels, num_classes, micro_batch=2, r=0.999, restore_bn_stats=True, pin=True):
    """Public end-to-end call: HOST images/labels -> EIC scores on the host.

    Protocol (the oracle follows the same one): for every micro-batch of
    `micro_batch` images (fixed by global index)  zero_grad -> loss = model(x, y, deepsup=True)
    -> backward -> one EIC step on the gradient of this step; BN runs in train mode on the
    micro-batch like the reference's training step (train.py:255-268), no optimizer step, running
    statistics restored afterwards.  With torch.distributed initialised, micro-batches are dealt
    round-robin to the ranks and each step's dgamma vector is averaged with one all-reduce before
    the sign gate (the reference gates on the DDP-averaged gradient, engine.py:66)."""
    ops.require_gpu()
    device = next(model.parameters()).device
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    world = torch.distributed.get_world_size() if dist_on else 1
    rank = torch.distributed.get_rank() if dist_on else 0
    scorer = ClassStatsScorer(model, num_classes, mode="bwd", r=r).attach()
    saved = None
    if restore_bn_stats:
        saved = [(m, m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone())
                 for m in model.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.running_mean is not None]
    was_training = model.training
    model.train()
    gamma = torch.cat([m.weight.detach().reshape(-1) for _, m in scorer.layers]).float()
    n = images.shape[0]
    n_steps = n // (micro_batch * world)
    h2d = d2h = 0
    try:
        for step in range(n_steps):
            lo = (step * world + rank) * micro_batch
            xb, yb = images[lo:lo + micro_batch], labels[lo:lo + micro_batch]
            if pin and not xb.is_pinned():
                xb, yb = xb.pin_memory(), yb.pin_memory()
            x = xb.to(device, non_blocking=True)
            y = yb.to(device, non_blocking=True)
            h2d += xb.numel() * xb.element_size() + yb.numel() * yb.element_size()
            scorer.zero_stats()
            scorer.set_labels(y)
            model.zero_grad(set_to_none=True)
            loss = model(x, y.long(), deepsup=True)
            loss = loss["loss"] if isinstance(loss, dict) else loss
            loss.backward()
            dgamma = scorer.dgamma()
            if world > 1:
                torch.distributed.all_reduce(dgamma)
                dgamma /= world
            scorer.eic_step(dgamma, gamma)
    finally:
        scorer.detach()
        model.train(was_training)
        if saved is not None:
            with torch.no_grad():
                for m, mean, var, nbt in saved:
                    m.running_mean.copy_(mean)
                    m.running_var.copy_(var)
                    m.num_batches_tracked.copy_(nbt)
    out = {"eic": {k: v.cpu() for k, v in scorer.eic_dict()["eic"].items()}}
    d2h += scorer.eic.numel() * 4
    out["_stats"] = dict(steps=n_steps, h2d_bytes=h2d, d2h_bytes=d2h, launches=scorer.launches)
    return out
