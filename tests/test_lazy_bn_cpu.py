"""lazy.PendingBN: the deferral that turns bn3 -> (+ shortcut) -> ReLU (networks/backbone/resnet.py:49-56) into one fused
call must be invisible to the program -- whatever a model does with the placeholder, values and gradients are those of the
eager modules.  Runs on CPU with a torch `run` standing in for the fused kernels (the protocol has no CUDA dependency);
the same checks against the real kernels are in tests/test_gpu_fused_scorer.py."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from dcfp_b200.lazy import PendingBN
from dcfp_b200.scorer import ClassStatsScorer, _FusedLayer
from dcfp_b200.workloads.segnets import Bottleneck


class _Patch:
    """bn.forward -> PendingBN whose `run` is plain torch; records how each placeholder was resolved."""

    def __init__(self, bn):
        self.bn, self.calls = bn, []
        bn.forward = self.forward

    def forward(self, x):
        bn = self.bn

        def run(residual, relu):
            self.calls.append((residual is not None, relu))
            y = F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.0, bn.eps)
            if residual is not None:
                y = y + residual
            return torch.relu(y) if relu else y

        return PendingBN.make(x, run, True, ("bn", False))

    def undo(self):
        self.bn.__dict__.pop("forward", None)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _make_block(seed=0, with_downsample=False):
    """A bottleneck of the workload nets (same tail as the reference's, resnet.py:49-56), tiny."""
    torch.manual_seed(seed)
    down = nn.Sequential(nn.Conv2d(16, 16, 1, bias=False), nn.BatchNorm2d(16)) if with_downsample else None
    blk = Bottleneck(16, 4, 1, 1, down)
    for m in blk.modules():
        if isinstance(m, nn.BatchNorm2d):
            nn.init.uniform_(m.weight, 0.5, 1.5)
            nn.init.uniform_(m.bias, -0.5, 0.5)
    return blk.train()


def _run_block(blk, x, patch_bn3):
    x = x.clone().requires_grad_(True)
    p = _Patch(blk.bn3) if patch_bn3 else None
    try:
        out = blk(_cl(x))
        out.square().sum().backward()
    finally:
        if p:
            p.undo()
    grads = [x.grad.clone()] + [q.grad.clone() for q in blk.parameters()]
    blk.zero_grad(set_to_none=True)
    return out.detach(), grads, (p.calls if p else None)


@pytest.mark.parametrize("with_downsample", [False, True])
def test_bottleneck_tail_is_one_fused_call_with_eager_results(with_downsample):
    blk = _make_block(with_downsample=with_downsample)
    x = torch.randn(2, 16, 6, 5)
    ref_out, ref_grads, _ = _run_block(blk, x, False)
    out, grads, calls = _run_block(blk, x, True)
    assert calls == [(True, True)], calls  # exactly one resolution: residual + ReLU together
    assert type(out) is torch.Tensor
    torch.testing.assert_close(out, ref_out, rtol=1e-6, atol=1e-6)
    for g, r in zip(grads, ref_grads):
        torch.testing.assert_close(g, r, rtol=1e-5, atol=1e-6)


def _pending(x, bn, calls):
    def run(residual, relu):
        calls.append((residual is not None, relu))
        y = F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.0, bn.eps)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y

    return PendingBN.make(x, run, True, ("bn", False))


def _eager(x, bn):
    return F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.0, bn.eps)


PROGRAMS = {
    # name: (program(p, r) -> tensor, expected list of resolutions)
    "add_relu_module": (lambda p, r: nn.ReLU(inplace=True)(p + r), [(True, True)]),
    "radd_relu": (lambda p, r: F.relu(r + p), [(True, True)]),
    "iadd_relu_": (lambda p, r: _iadd(p, r).relu_(), [(True, True)]),
    "torch_add_alpha1": (lambda p, r: torch.relu(torch.add(p, r, alpha=1)), [(True, True)]),
    "relu_without_residual": (lambda p, r: F.relu(p, inplace=True) + r, [(False, True)]),
    "alpha2_materializes": (lambda p, r: torch.relu(torch.add(p, r, alpha=2)), [(False, False)]),
    "mul_materializes": (lambda p, r: (p * 2.0 + r).relu(), [(False, False)]),
    "scalar_add": (lambda p, r: F.relu(p + 1.0) + r, [(False, False)]),
    "sum_used_twice": (lambda p, r: _twice(p, r), [(True, True), (False, False)]),
    "bn_output_used_too": (lambda p, r: _both(p, r), [(False, False)]),
    "cat": (lambda p, r: torch.cat([p, r], 1).relu()[:, :p.shape[1]], [(False, False)]),
    "strided_residual": (lambda p, r: F.relu(p + r.contiguous()), [(False, False)]),
    "broadcast_residual": (lambda p, r: F.relu(p + r[:, :, :1, :1]), [(False, False)]),
    "double_add": (lambda p, r: F.relu((p + r) + r), [(False, False)]),
    "conv_consumer": (lambda p, r: F.conv2d(p, torch.full((8, 8, 1, 1), 0.125)) + r, [(False, False)]),
    "sub_materializes": (lambda p, r: F.relu(p - r), [(False, False)]),
    "dtype_conversion": (lambda p, r: F.relu(p.double() + r.double()).float(), [(False, False)]),
    "type_conversion": (lambda p, r: F.relu(p.type(torch.float64) + r.double()).float(), [(False, False)]),
    "view_then_add": (lambda p, r: F.relu(p.view_as(r) + r), [(False, False)]),
    "tail_plus_detached_bn_output": (lambda p, r: F.relu(p + r) + p.detach(), [(True, True), (False, False)]),
    "metadata_only": (lambda p, r: F.relu(p + r) * float(p.shape[1] + p.dim() + p.size(0) + int(p.is_contiguous(memory_format=torch.channels_last))),
                      [(True, True)]),
}


def _iadd(p, r):
    p += r
    return p


def _twice(p, r):
    s = p + r
    return F.relu(s) + s  # non-inplace ReLU, then the un-rectified sum again


def _both(p, r):
    z = p * 1.0  # the BN output itself is consumed first ...
    return F.relu(p + r) + z  # ... so the tail runs on the materialised value


@pytest.mark.parametrize("name", sorted(PROGRAMS))
def test_any_use_of_the_placeholder_gives_eager_values_and_gradients(name):
    prog, expected = PROGRAMS[name]
    torch.manual_seed(1)
    bn = nn.BatchNorm2d(8)
    nn.init.uniform_(bn.weight, 0.5, 1.5)
    nn.init.uniform_(bn.bias, -0.5, 0.5)
    x0, r0 = _cl(torch.randn(2, 8, 4, 3)), _cl(torch.randn(2, 8, 4, 3))

    def grads(use_pending):
        x, r = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        calls = []
        xin, rin = x * 1.0, r * 1.0  # non-leaf, as inside a network
        p = _pending(xin, bn, calls) if use_pending else _eager(xin, bn)
        out = prog(p, rin)
        assert type(out) is torch.Tensor
        (out * torch.arange(out.numel(), dtype=out.dtype).view_as(out)).sum().backward()
        g = [x.grad.clone(), r.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()]
        bn.zero_grad(set_to_none=True)
        return out.detach(), g, calls

    ref_out, ref_g, _ = grads(False)
    out, g, calls = grads(True)
    assert calls == expected, (name, calls)
    torch.testing.assert_close(out, ref_out, rtol=1e-6, atol=1e-6)
    for a, b in zip(g, ref_g):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


def test_placeholder_answers_metadata_without_resolving():
    bn = nn.BatchNorm2d(8)
    calls = []
    x = _cl(torch.randn(2, 8, 4, 3))
    p = _pending(x, bn, calls)
    assert p.shape == x.shape and p.dtype == x.dtype and p.device == x.device and p.dim() == 4 and p.numel() == x.numel()
    assert p.is_contiguous(memory_format=torch.channels_last) and p.stride() == x.stride() and p.requires_grad
    assert isinstance(p, torch.Tensor) and calls == []
    q = p + _cl(torch.randn(2, 8, 4, 3))
    assert isinstance(q, PendingBN) and q.shape == x.shape and calls == []
    assert getattr(q, "_dcfp_bn") == ("bn", False)
    # ... but `type` WITH an argument is a conversion of the values, not a metadata read: it must see the layer's output
    assert p.type() == "torch.FloatTensor" and calls == []
    z = p.type(torch.float64)
    assert calls == [(False, False)] and type(z) is torch.Tensor
    torch.testing.assert_close(z, _eager(x, bn).double())


def test_residual_tail_is_learned_from_the_autograd_graph():
    """scorer._learn_residual_tail reads `relu input = AddBackward0(fused BN node, ...)`; a stand-in Function carries the same
    ctx attributes the fused one sets (layer, relu, has_res)."""

    class FakeScorer:
        _add_relu_after = {}

    sc = FakeScorer()
    layer = _FusedLayer(sc, "blk.bn3", None, None, None, 0)
    other = _FusedLayer(sc, "blk.downsample.1", None, None, None, 1)

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, lay, relu):
            ctx.layer, ctx.relu, ctx.has_res = lay, relu, False
            return x * 2

        @staticmethod
        def backward(ctx, g):
            return g * 2, None, None

    x = torch.randn(3, requires_grad=True)
    learn = ClassStatsScorer._learn_residual_tail
    learn(sc, Fn.apply(x, layer, True) + x)  # that BN is already fused with a ReLU of its own: not a tail
    learn(sc, Fn.apply(x, layer, False) * x)  # not an add
    learn(sc, torch.add(Fn.apply(x, layer, False), x, alpha=2))
    learn(sc, x + 1.0)
    learn(sc, torch.randn(3))  # no graph
    assert sc._add_relu_after == {}
    learn(sc, Fn.apply(x, layer, False) + Fn.apply(x, other, False))  # bn3 + downsample BN: the first operand defers
    assert sc._add_relu_after == {"blk.bn3": True}
    sc._add_relu_after.clear()
    learn(sc, x.detach() + Fn.apply(x, other, False))
    assert sc._add_relu_after == {"blk.downsample.1": True}


@pytest.mark.ref
def test_reference_bottleneck_tail_is_one_fused_call():
    """The unmodified reference block (networks/backbone/resnet.py:20-57): same protocol, same results."""
    from oracle import ref_compat
    ref = ref_compat.load_reference()
    import networks.backbone.resnet as ref_resnet
    assert ref_resnet.__file__.startswith(ref_compat.REF_ROOT)
    del ref
    torch.manual_seed(0)
    down = nn.Sequential(nn.Conv2d(16, 16, 1, bias=False), nn.BatchNorm2d(16))
    for blk in (ref_resnet.Bottleneck(16, 4), ref_resnet.Bottleneck(16, 4, downsample=down)):
        blk.train()
        x = torch.randn(2, 16, 6, 5)
        ref_out, ref_grads, _ = _run_block(blk, x, False)
        out, grads, calls = _run_block(blk, x, True)
        assert calls == [(True, True)], calls
        torch.testing.assert_close(out, ref_out, rtol=1e-6, atol=1e-6)
        for g, r in zip(grads, ref_grads):
            torch.testing.assert_close(g, r, rtol=1e-5, atol=1e-6)


# ---- the backward half: shortcut gradients deposited on the producing tail's node (lazy.deposit_shortcut_grad) ---------------
class _Tail(torch.autograd.Function):
    """y = relu(a * x + res) with scorer._FusedBN's deposit protocol (a toy stand-in for bn3 + shortcut + ReLU)."""

    @staticmethod
    def forward(ctx, x, res, a, res_node, ledger, log):
        y = torch.relu(a * x + res)
        ctx.save_for_backward(y)
        ctx.a, ctx.ledger, ctx.log, ctx.has_res = a, ledger, log, True
        ctx.res_node = res_node if ctx.needs_input_grad[1] else None
        return y

    @staticmethod
    def backward(ctx, dy):
        from dcfp_b200.lazy import deposit_shortcut_grad, take_shortcut_grad
        (y,) = ctx.saved_tensors
        dep = take_shortcut_grad(ctx, ctx.ledger)
        ctx.log.append("fused" if dep is not None else "plain")
        dz = (dy if dep is None else dy + dep) * (y > 0)
        gres = None
        if ctx.res_node is not None:
            deposit_shortcut_grad(ctx.res_node, dz, ctx.ledger)
        elif ctx.needs_input_grad[1]:
            gres = dz
        return ctx.a * dz, gres, None, None, None, None


def _tower(x0, ws, deposit):
    ledger, log = [0], []
    t = x0 * 1.0
    for w in ws:
        branch = torch.sin(t) * w  # the conv1 .. bn3 path of a block
        if deposit is None:
            t = torch.relu(0.7 * branch + t)
        else:
            node = t.grad_fn if (deposit and isinstance(t.grad_fn, _Tail._backward_cls)) else None
            t = _Tail.apply(branch, t, 0.7, node, ledger, log)
    return t, ledger, log


def test_deposited_shortcut_gradients_equal_autograd_accumulation():
    torch.manual_seed(0)
    x0 = torch.randn(4, 6, dtype=torch.float64)
    ws0 = [torch.randn(6, dtype=torch.float64) for _ in range(5)]
    weight = torch.randn(4, 6, dtype=torch.float64)
    res = {}
    for mode in (None, False, True):
        x = x0.clone().requires_grad_(True)
        ws = [w.clone().requires_grad_(True) for w in ws0]
        out, ledger, log = _tower(x, ws, mode)
        (out * weight).sum().backward()
        res[mode] = [x.grad] + [w.grad for w in ws]
        assert ledger[0] == 0
        if mode:  # backward order: the last block has no consumer that deposits; the four before it fold a deposit into their gate
            assert log == ["plain"] + ["fused"] * 4, log
        elif mode is False:
            assert log == ["plain"] * 5
    for a, b, c in zip(res[None], res[False], res[True]):
        torch.testing.assert_close(b, a, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(c, a, rtol=1e-12, atol=1e-12)


def test_unconsumed_deposit_is_detected():
    """A tail whose output feeds ONLY a shortcut: autograd may never run (or run with zeros) the producer -- the ledger says so."""
    from dcfp_b200.scorer import ClassStatsScorer

    class Sc:
        _deposits = [0]
    ledger, log = Sc._deposits, []
    x = torch.randn(3, dtype=torch.float64, requires_grad=True)
    z = torch.randn(3, dtype=torch.float64, requires_grad=True)
    t1 = _Tail.apply(x * 1.0, x * 0.5, 0.7, None, ledger, log)
    t2 = _Tail.apply(z * 1.0, t1.detach().requires_grad_(True), 0.7, t1.grad_fn, ledger, log)  # graph cut: t1's node never runs
    t2.sum().backward()
    assert ledger[0] == 1
    with pytest.raises(RuntimeError, match="never consumed"):
        ClassStatsScorer.check_deposits(Sc)
    assert Sc._deposits[0] == 0
