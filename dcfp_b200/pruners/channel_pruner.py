"""Host-side mirror of the reference's channel pruner (pruners/channel_pruner.py) -- same names,
argument meaning and results, new implementation.

What stays Python (topology work, no data parallelism -- SURVEY.md 0.3):
  * the autograd-graph analysis that yields `norm_conv_links`, `node2parents`, channel groups and
    spaces (reference :139-255, :423-476, :501-737).  Re-implemented as two ITERATIVE walks that
    reproduce the reference's discovery order without materialising or deep-copying paths, and
    cached per architecture so the 0.5 -> 1.0 `global_percent` loop of prune.py traces once;
  * mask propagation through groups / concats (:750-819) on C-float vectors.
What runs on the GPU through the C ABI (no CPU fallback):
  * bias compensation  offset = W.sum((2,3)) @ relu((1-in_mask)*beta)     (:873-905) -> dcfp_bias_comp
  * the weight / bias / running-stat slicing of `deploy_subnet`              (:907-948) -> dcfp_channel_gather_grouped
"""
import copy
from collections import OrderedDict
from types import MethodType

import torch
import torch.nn as nn
from torch.nn.modules import GroupNorm
from torch.nn.modules.batchnorm import _BatchNorm
from torch.nn.modules.instancenorm import _InstanceNorm

from .. import ops

# autograd node-name prefixes (reference :12-23); torch >= 1.11 calls every conv `ConvolutionBackward0`
CONV = ("ConvolutionBackward", "ThnnConv2DBackward", "CudnnConvolutionBackward", "MkldnnConvolutionBackward",
        "SlowConvDilated2DBackward")
FC = ("ThAddmmBackward", "AddmmBackward", "MmBackward")
BN = ("ThnnBatchNormBackward", "CudnnBatchNormBackward", "NativeBatchNormBackward")
GN = ("NativeGroupNormBackward",)
CONCAT = ("CatBackward",)
NORM = BN + GN


def _kind(fn):
    name = type(fn).__name__
    if name.startswith(CONV):
        return "conv"
    if name.startswith(FC):
        return "fc"
    if name.startswith(CONCAT):
        return "cat"
    return None


class _OrderedSet:
    """Insertion-ordered set (stands in for the un-vendored `ordered_set` the reference imports)."""

    def __init__(self, items=()):
        self._d = dict.fromkeys(items)

    def add(self, x):
        self._d.setdefault(x, None)

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, x):
        return x in self._d

    def __getitem__(self, i):
        return list(self._d)[i]

    def __repr__(self):
        return "OrderedSet(%r)" % (list(self._d),)


def init_pruned_model(supernet, channel_cfg):
    """Shape-only prefix slicing of a FRESH model to `channel_cfg` sizes before load_state_dict
    (reference :29-74; consumers prune.py:109, train.py:202, evaluate.py:289).  No data dependence,
    so it stays a host-side view operation."""
    for name, module in supernet.named_modules():
        cfg = channel_cfg.get(name)
        if cfg is None:
            continue
        requires_grad = module.weight.requires_grad
        out_channels = cfg["out_channels"]
        weight = module.weight[:out_channels]
        for attr in ("out_channels", "out_features", "num_features"):
            if hasattr(module, attr):
                setattr(module, attr, out_channels)
        if "in_channels" in cfg:
            in_channels = cfg["in_channels"]
            weight = weight[:, :in_channels]
            for attr in ("in_channels", "in_features"):
                if hasattr(module, attr):
                    setattr(module, attr, in_channels)
            if getattr(module, "groups", in_channels) > 1:
                module.groups = in_channels
        module.weight = nn.Parameter(weight.data.contiguous())
        module.weight.requires_grad = requires_grad
        if getattr(module, "bias", None) is not None:
            module.bias = nn.Parameter(module.bias[:out_channels].data.contiguous())
            module.bias.requires_grad = requires_grad
        if hasattr(module, "running_mean"):
            module.running_mean = module.running_mean[:out_channels].contiguous()
        if hasattr(module, "running_var"):
            module.running_var = module.running_var[:out_channels].contiguous()


def _structural_clone(model):
    """A copy of the module tree that SHARES every parameter / buffer with `model`.

    The reference deep-copies the whole network (1.6-2.6 s, 184-282 MB) only to hang mask buffers
    on it (:968,973); sharing storage gives the same container for the masks at no cost."""
    memo = {}
    for t in list(model.parameters()) + list(model.buffers()):
        memo[id(t)] = t
    return copy.deepcopy(model, memo)


_TOPOLOGY_CACHE = {}


class ChannelPruner:
    """Structure pruner base class (reference :108-990)."""

    #: pseudo input of the tracing forward (reference :191)
    trace_input_size = (2, 3, 224, 224)

    def __init__(self, except_start_keys=None, **kwards):
        self.except_start_keys = list() if except_start_keys is None else except_start_keys
        self.end_nodes = []

    # ------------------------------------------------------------------ graph analysis
    def add_pruning_attrs(self, module):
        """in_mask / out_mask buffers and a masked forward (reference :375-421)."""
        if isinstance(module, nn.Conv2d):
            module.register_buffer("in_mask", module.weight.new_ones((1, module.in_channels, 1, 1)))
            module.register_buffer("out_mask", module.weight.new_ones((1, module.out_channels, 1, 1)))
            inner = module.forward

            def conv_forward(self, feature):
                return inner(feature * self.in_mask)

            module.forward = MethodType(conv_forward, module)
        if isinstance(module, nn.Linear):
            module.register_buffer("in_mask", module.weight.new_ones((1, module.in_features)))
            module.register_buffer("out_mask", module.weight.new_ones((1, module.out_features)))
            inner_fc = module.forward

            def fc_forward(self, feature):
                if len(self.in_mask.shape) != len(self.out_mask.shape):
                    self.in_mask = self.in_mask.reshape(self.in_mask.shape[:2])
                return inner_fc(feature * self.in_mask)

            module.forward = MethodType(fc_forward, module)
        if isinstance(module, (_BatchNorm, _InstanceNorm, GroupNorm)):
            module.register_buffer("out_mask", module.weight.new_ones((1, len(module.weight), 1, 1)))

    def _trace(self, supernet, weighted):
        """Runs the tracing forward and both graph walks; returns plain-Python topology."""
        var2name = {id(m.weight): n for n, m in weighted.items()}
        calls = dict.fromkeys(weighted, 0)
        handles = []
        for n, m in weighted.items():
            handles.append(m.register_forward_hook(lambda mod, i, o, _n=n: calls.__setitem__(_n, calls[_n] + 1)))
        grad_flags = [(p, p.requires_grad) for p in supernet.parameters()]
        for p, _ in grad_flags:
            p.requires_grad = True
        modes = [(m, m.training) for m in supernet.modules()]
        supernet.eval()
        try:
            ref_param = next(supernet.parameters())
            pseudo_img = torch.randn(*self.trace_input_size).to(device=ref_param.device, dtype=ref_param.dtype)
            with torch.enable_grad():  # a caller's no_grad() must not yield an empty (and then cached) topology
                out = supernet.forward(pseudo_img, deepsup=True)
            if isinstance(out, (list, tuple)):
                loss = sum(o.sum() for o in out)
            elif isinstance(out, dict):
                loss = sum(out[k].sum() for k in out)
            else:
                loss = out.sum()
        finally:
            for h in handles:
                h.remove()
            for p, flag in grad_flags:
                p.requires_grad = flag
            for m, mode in modes:
                m.training = mode
        shared = [n for n, c in calls.items() if c >= 2]
        node2parents = self._walk_non_pass(loss.grad_fn, var2name, set(shared))
        links = self._walk_norm_links(loss.grad_fn, var2name)
        return dict(shared=shared, node2parents=[(k, list(v)) for k, v in node2parents.items()], links=list(links.items()))

    def _walk_non_pass(self, root, var2name, shared):
        """node -> ordered parents over conv / linear / concat nodes (reference :501-520, :621-737,
        :423-448).  The reference enumerates every root-to-leaf path (deep-copying it) and then scans
        consecutive pairs; only the not-yet-scanned suffix of a path can add anything, so the same
        insertion order falls out of one walk with a live path stack."""
        node2parents = OrderedDict()
        visited = {}
        path, done = [], 0  # `done` = leading pairs of `path` already recorded
        keep_alive, cat_names = [], {}
        end_nodes = self.end_nodes

        def emit():
            nonlocal done
            if not path:
                return
            for i in range(done, len(path) - 1):
                node, parent = path[i], path[i + 1]
                if parent in end_nodes:
                    continue
                node2parents.setdefault(node, _OrderedSet()).add(parent)
            node2parents.setdefault(path[-1], _OrderedSet())
            done = len(path) - 1

        stack = [("visit", root)]
        while stack:
            op, arg = stack.pop()
            if op == "push":
                path.append(arg)
                continue
            if op == "pop":
                path.pop()
                done = min(done, max(len(path) - 1, 0))
                continue
            fn = arg
            if fn is None:  # reached the network input
                emit()
                continue
            kind = _kind(fn)
            if kind in ("conv", "fc"):
                if kind == "conv":
                    var, parent = fn.next_functions[1][0].variable, fn.next_functions[0][0]
                else:
                    var, parent = fn.next_functions[-1][0].next_functions[0][0].variable, fn.next_functions[1][0]
                name = var2name[id(var)]
                path.append(name)
                stack.append(("pop", None))
                if visited.get(name) and name not in shared:
                    emit()
                else:
                    visited[name] = True
                    stack.append(("visit", parent))
            elif kind == "cat":
                if id(fn) not in cat_names:
                    keep_alive.append(fn)
                    cat_names[id(fn)] = "concat_%d" % len(cat_names)
                name = cat_names[id(fn)]
                path.append(name)
                stack.append(("pop", None))
                if visited.get(name):
                    emit()
                else:
                    visited[name] = True
                    parents = fn.next_functions
                    for i in range(len(parents) - 1, -1, -1):
                        stack.append(("pop", None))
                        stack.append(("visit", parents[i][0]))
                        stack.append(("push", "%s_item_%d" % (name, i)))
            else:
                for nxt in reversed(fn.next_functions):
                    stack.append(("visit", nxt[0]))
        return node2parents

    def _walk_norm_links(self, root, var2name):
        """{norm name: name of the conv feeding it}, in discovery order from the loss (reference :522-614)."""

        def leaf_var(fn):
            while type(fn).__name__ != "AccumulateGrad":
                fn = fn.next_functions[0][0]
            return fn.variable

        links = OrderedDict()
        stack = [root]
        while stack:
            fn = stack.pop()
            if fn is None:
                continue
            if type(fn).__name__.startswith(NORM):
                conv_fn = fn.next_functions[0][0]
                while not type(conv_fn).__name__.startswith(CONV):
                    if conv_fn is None or not conv_fn.next_functions:
                        raise AttributeError("normalisation layer without a convolution above it in the autograd graph")
                    conv_fn = conv_fn.next_functions[0][0]
                conv_name = var2name[id(leaf_var(conv_fn.next_functions[1][0]))]
                bn_name = var2name[id(leaf_var(fn.next_functions[1][0]))]
                if bn_name not in links:
                    links[bn_name] = conv_name
                    stack.append(conv_fn)
            else:
                for nxt in reversed(fn.next_functions):
                    stack.append(nxt[0])
        return links

    def prepare_from_supernet(self, supernet):
        """Attach masks and derive the channel topology (reference :139-255)."""
        name2module, module2name = OrderedDict(), OrderedDict()
        for name, module in supernet.named_modules():
            if isinstance(module, nn.GroupNorm):
                raise NotImplementedError("GroupNorm tracing is not supported (the reference path raises NameError here, :161)")
            if hasattr(module, "weight"):
                name2module[name] = module
                module2name[module] = name
        self.name2module, self.module2name = name2module, module2name

        cls = type(supernet)
        key = (cls.__module__, cls.__qualname__, id(getattr(cls.forward, "__code__", None)), tuple(self.trace_input_size),
               tuple(self.end_nodes), tuple((n, type(m).__name__, tuple(m.weight.shape)) for n, m in name2module.items()),
               tuple(sorted((k, repr(v)) for k, v in vars(supernet).items() if isinstance(v, (bool, int, str)) and not k.startswith("_"))))
        topo = _TOPOLOGY_CACHE.get(key)
        if topo is None:
            topo = self._trace(supernet, name2module)
            if topo["links"]:  # never cache a trace that found nothing (the next call re-traces, like the reference)
                _TOPOLOGY_CACHE[key] = topo
        else:
            # the reference traces every time and its pseudo image comes from the GLOBAL torch generator (:191); draw
            # the same numbers so that RNG-dependent callers (RandomChannelPruner) see the reference's stream
            torch.randn(*self.trace_input_size)
        for module in name2module.values():
            self.add_pruning_attrs(module)

        self.shared_module = list(topo["shared"])
        self.norm_conv_links = dict(topo["links"])
        self.conv_norm_links = {conv: norm for norm, conv in self.norm_conv_links.items()}
        self.node2parents = OrderedDict((k, _OrderedSet(v)) for k, v in topo["node2parents"])
        self.same_out_channel_groups = self.make_same_out_channel_groups(self.node2parents, name2module)
        self.module2group = {m: g for g, members in self.same_out_channel_groups.items() for m in members}
        self.modules_have_ancest = [n for n, p in self.node2parents.items() if n in name2module and len(p) > 0]
        self.modules_have_child = _OrderedSet(p for parents in self.node2parents.values() for p in parents if p in name2module)
        self.channel_spaces = self.build_channel_spaces(name2module)

    def make_same_out_channel_groups(self, node2parents, name2module):
        """Modules that feed a common child must keep identical out-channels (reference :293-373).
        Nodes are taken in discovery order; a node joins the FIRST group its parents intersect."""
        children, parents_of = OrderedDict(), OrderedDict()
        for node, parents in node2parents.items():
            is_module = node in name2module
            if "concat" in node and not is_module:
                if "item" not in node:
                    continue
            elif "chunk" in node and not is_module:
                continue
            pset = list(parents)
            if is_module and isinstance(name2module[node], nn.Conv2d) and name2module[node].in_channels == name2module[node].groups:
                if node not in pset:  # depth-wise conv shares its parent's channels
                    pset.append(node)
                    parents.add(node)
            for gid, members in parents_of.items():
                if any(p in members for p in pset):
                    children[gid].append(node)
                    parents_of[gid] = pset + [m for m in members if m not in pset]
                    break
            else:
                gid = len(parents_of)
                children[gid] = [node]
                parents_of[gid] = pset
        groups = OrderedDict()
        for members in parents_of.values():
            if len(members) > 1:
                groups["group_%d" % len(groups)] = members
        return groups

    def build_channel_spaces(self, name2module):
        """space id -> out_mask of its first member (reference :450-476)."""
        spaces = OrderedDict()
        for name in self.modules_have_child:
            spaces.setdefault(self.module2group.get(name, name), name2module[name].out_mask)
        return spaces

    def get_space_id(self, module_name):
        """(reference :257-291)"""
        if "concat" in module_name and module_name not in self.name2module:
            if "item" in module_name:
                return self.get_space_id(self.node2parents[module_name][0])
            return dict(concat=[self.get_space_id(p) for p in self.node2parents[module_name]])
        if module_name not in self.modules_have_child:
            return None
        return self.module2group.get(module_name, module_name)

    # ------------------------------------------------------------------ masks
    def gen_channel_mask(self):
        pass

    def get_channel_mask(self, space_id, out_mask):
        """(reference :750-761) group = union of its members' masks, concat = concatenation."""
        if isinstance(space_id, dict):
            return torch.cat([self.get_channel_mask(p, out_mask) for p in space_id["concat"]])
        if space_id in self.same_out_channel_groups:
            mask = torch.zeros_like(out_mask)
            for member in self.same_out_channel_groups[space_id]:
                mask = mask + self.get_channel_mask(member, out_mask)
            return torch.clamp(mask, 0, 1)
        return self.name2module[space_id].out_mask

    def sample_subnet(self):
        return {sid: self.get_channel_mask(sid, m) for sid, m in self.channel_spaces.items()}

    def _parent_tensor(self, name, table):
        """value of `table` for the first parent's space (concat -> cat along channels)."""
        space_id = self.get_space_id(self.node2parents[name][0])
        if isinstance(space_id, dict):
            return torch.cat([table[p] for p in space_id["concat"]], dim=1)
        return table[space_id]

    def set_subnet(self, subnet_dict):
        """(reference :775-819)"""
        for name in self.modules_have_child:
            module = self.name2module[name]
            module.out_mask = subnet_dict[self.get_space_id(name)].to(module.out_mask.device)
        for norm, conv in self.norm_conv_links.items():
            space_id = self.get_space_id(conv)
            if space_id is not None:  # None: the conv in front is an output layer, never out-pruned
                module = self.name2module[norm]
                module.out_mask = subnet_dict[space_id].to(module.out_mask.device)
        for name in self.modules_have_ancest:
            module = self.name2module[name]
            module.in_mask = self._parent_tensor(name, subnet_dict).to(module.in_mask.device)

    def export_subnet(self):
        """(reference :821-842)"""
        channel_cfg = dict()
        for name, module in self.name2module.items():
            cfg = channel_cfg[name] = dict()
            for side in ("in", "out"):
                mask = getattr(module, side + "_mask", None)
                if mask is not None:
                    cfg[side + "_channels"] = int(mask.sum())
                    cfg["raw_" + side + "_channels"] = int(mask.numel())
                    cfg[side + "_mask"] = mask.cpu().numpy()
        return channel_cfg

    # ------------------------------------------------------------------ bias compensation
    def get_space_bias(self, space_id, out_mask):
        """(reference :844-859)"""
        if isinstance(space_id, dict):
            return torch.cat([self.get_space_bias(p, out_mask) for p in space_id["concat"]])
        if space_id in self.same_out_channel_groups:
            bias = torch.zeros_like(out_mask)
            for member in self.same_out_channel_groups[space_id]:
                bias = bias + self.get_space_bias(member, out_mask)
            return bias
        if space_id in self.conv_norm_links:
            return self.name2module[self.conv_norm_links[space_id]].bias.reshape(out_mask.shape)
        return torch.zeros_like(out_mask)

    def get_subnet_bias(self):
        return {sid: self.get_space_bias(sid, m) for sid, m in self.channel_spaces.items()}

    def resize_subnet_bias(self, supernet, bias_dict):
        """Fold the constant activation relu(beta) of every pruned input channel into the consumer
        (reference :873-905): offset = W.sum((2,3)) @ relu((1 - in_mask) * beta_parent), subtracted
        from the next BN's running_mean (or added to the conv bias).  The reduce-GEMV runs on the GPU;
        when no pruned channel has beta > 0 the offset is exactly zero and nothing is launched."""
        for name, module in supernet.named_modules():
            if name not in self.modules_have_ancest:
                continue
            sub_module = self.name2module[name]
            bias = self._parent_tensor(name, bias_dict)
            activation = torch.relu((1 - sub_module.in_mask) * bias.detach())
            if bool((activation != 0).any()):
                ops.require_gpu()
                device = ops.device()
                w = module.weight.data
                offset = ops.bias_comp(w.to(device).contiguous(), activation.reshape(-1).to(device=device, dtype=torch.float32))
                offset = offset.to(w.device)
            else:
                offset = torch.zeros(module.weight.shape[0], dtype=module.weight.dtype, device=module.weight.device)
            if name in self.conv_norm_links:
                supernet.get_submodule(self.conv_norm_links[name]).running_mean.data.sub_(offset)
            elif hasattr(sub_module, "bias"):
                module.bias.data.add_(offset)
            else:
                module.bias = nn.Parameter(offset)

    # ------------------------------------------------------------------ deploy (K3)
    def deploy_subnet(self, supernet, channel_cfg):
        """Slice every weighted module to its kept channels (reference :907-948) -- ONE grouped gather
        launch over all weights, biases and running statistics of the model."""
        ops.require_gpu()
        device = ops.device()
        jobs = []  # (module, attr kind, source tensor, out mask id, in mask id)
        masks, mask_ids = [], {}

        def mask_id(mask):
            flat = mask.reshape(-1)
            key = id(mask)
            if key not in mask_ids:
                mask_ids[key] = len(masks)
                masks.append(torch.nonzero(flat == 1).reshape(-1).to(torch.int32).cpu())
            return mask_ids[key]

        for name, module in supernet.named_modules():
            if name not in channel_cfg:
                continue
            sub_module = self.name2module[name]
            out_id = mask_id(sub_module.out_mask)
            in_id = mask_id(sub_module.in_mask) if hasattr(sub_module, "in_mask") else None
            out_channels = int(sub_module.out_mask.sum())
            for attr in ("out_channels", "out_features", "num_features"):
                if hasattr(module, attr):
                    setattr(module, attr, out_channels)
            if in_id is not None:
                in_channels = int(sub_module.in_mask.sum())
                for attr in ("in_channels", "in_features"):
                    if hasattr(module, attr):
                        setattr(module, attr, in_channels)
                if getattr(module, "groups", in_channels) > 1:
                    module.groups = in_channels
            jobs.append((module, "weight", sub_module.weight.data, out_id, in_id, sub_module.weight.requires_grad))
            if getattr(module, "bias", None) is not None:
                jobs.append((module, "bias", module.bias.data, out_id, None, sub_module.weight.requires_grad))
            if hasattr(module, "running_mean"):
                jobs.append((module, "running_mean", module.running_mean, out_id, None, None))
            if hasattr(module, "running_var"):
                jobs.append((module, "running_var", module.running_var, out_id, None, None))
        if not jobs:
            return
        # all index lists travel in one host->device copy; per-mask views are slices of it
        sizes = [m.numel() for m in masks]
        flat_idx = torch.cat(masks).to(device) if sum(sizes) else torch.empty(0, dtype=torch.int32, device=device)
        views, pos = [], 0
        for n in sizes:
            views.append(flat_idx[pos:pos + n])
            pos += n
        by_size = {}
        for j, job in enumerate(jobs):
            by_size.setdefault(job[2].element_size(), []).append(j)
        results = [None] * len(jobs)
        for idxs in by_size.values():
            srcs = [jobs[j][2].to(device).contiguous() for j in idxs]
            outs = ops.channel_gather_grouped(srcs, [views[jobs[j][3]] for j in idxs],
                                              [None if jobs[j][4] is None else views[jobs[j][4]] for j in idxs])
            for j, out in zip(idxs, outs):
                results[j] = out
        for (module, attr, src, _o, _i, requires_grad), out in zip(jobs, results):
            out = out.to(src.device)
            if attr in ("weight", "bias"):
                param = nn.Parameter(out)
                param.requires_grad = requires_grad
                setattr(module, attr, param)
            else:
                setattr(module, attr, out)

    # ------------------------------------------------------------------ driver
    def get_except_layers(self, supernet):
        """(reference :950-965) exact keys pull in their BN/conv partner, then prefix match."""
        keys = []
        for key in self.except_start_keys:
            keys.append(key)
            if key in self.norm_conv_links:
                keys.append(self.norm_conv_links[key])
            elif key in self.conv_norm_links:
                keys.append(self.conv_norm_links[key])
        self.except_layers = [name for name, module in supernet.named_modules()
                              if hasattr(module, "weight") and any(name.startswith(k) for k in keys)]

    def prune_model(self, supernet, except_start_keys=None):
        """Returns (supernet sliced in place, channel_cfg) -- reference :967-990."""
        ops.require_gpu()
        model_copy = _structural_clone(supernet)
        self.end_nodes = getattr(model_copy, "end_nodes", [])
        self.prepare_from_supernet(model_copy)
        if hasattr(model_copy, "ignore_prune_layer"):
            self.except_start_keys = self.except_start_keys + model_copy.ignore_prune_layer
        if except_start_keys:
            self.except_start_keys = self.except_start_keys + except_start_keys
        self.get_except_layers(model_copy)

        self.gen_channel_mask()
        self.set_subnet(self.sample_subnet())
        self.resize_subnet_bias(supernet, self.get_subnet_bias())
        channel_cfg = self.export_subnet()
        self.deploy_subnet(supernet, channel_cfg)
        return supernet, channel_cfg
