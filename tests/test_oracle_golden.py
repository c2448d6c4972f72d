"""The oracle restatements against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) -- runs on CPU, no reference tree needed."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import class_stats_ref, eic_ref, scoring_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("r", [0.999, 0.99])
def test_eic_restatement_matches_reference_bits(r):
    """oracle/eic_ref.py vs dcfp_pruning.step of the reference (pruners/dcfp_pruner.py:15-20): bit-exact,
    including the int-0 first step, zero / NaN / inf / underflowing gradients and negative gammas."""
    z = np.load(os.path.join(GOLDEN, "eic_steps.npz"))
    sizes, steps = z["sizes"], int(z["steps"])
    tag = str(r).replace(".", "p")
    for i in range(len(sizes)):
        if i == 3:  # ignore_prune_layer in the fixture
            continue
        eic = 0
        for t in range(steps):
            eic = eic_ref.eic_step(eic, z["grad_%d_%d" % (t, i)], z["gamma_%d" % i], r)
            exp = z["eic_r%s_%d_%d" % (tag, t, i)]
            same = (_bits(eic) == _bits(exp)) | (np.isnan(eic) & np.isnan(exp))
            assert same.all(), "layer %d step %d: %d mismatching channels" % (i, t, (~same).sum())


@pytest.mark.parametrize("shape", [((512, 1024), (64, 128)), ((512, 512), (6, 6)), ((512, 512), (3, 3)), ((769, 769), (97, 97)),
                                   ((512, 1024), (1, 1)), ((100, 37), (13, 9)), ((64, 64), (128, 128))])
def test_nearest_labels_matches_interpolate(shape):
    """index-math restatement of legacy `nearest` == F.interpolate(mode='nearest')."""
    (h0, w0), (h, w) = shape
    g = torch.Generator().manual_seed(h0 * 7 + w)
    lab = torch.randint(0, 200, (2, h0, w0), generator=g)
    exp = F.interpolate(lab[:, None].float(), size=(h, w), mode="nearest")[:, 0].long()
    assert torch.equal(class_stats_ref.nearest_labels(lab, h, w), exp)


def test_class_stats_bwd_sums_to_autograd_dgamma():
    """sum_k S1[k, c] of the backward functor == bn.weight.grad of torch's autograd (the quantity
    pruners/dcfp_pruner.py:18 reads) -- pins the oracle's `bwd` value functor to the reference path."""
    torch.manual_seed(0)
    K, N, C, h, w = 7, 2, 24, 12, 20
    bn = torch.nn.BatchNorm2d(C).double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    x = torch.randn(N, C, h, w, dtype=torch.float64, requires_grad=True)
    y = bn(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    label = torch.randint(0, K, (N, 4 * h, 4 * w))
    label[0, :9, :9] = 255
    mean = x.detach().mean(dim=(0, 2, 3))
    invstd = torch.rsqrt(x.detach().var(dim=(0, 2, 3), unbiased=False) + bn.eps)
    v = class_stats_ref.functor_bwd(x.detach(), dy, invstd, -mean * invstd)
    # with an ignore label the class sum misses those pixels: compare against the masked gradient
    lab = class_stats_ref.nearest_labels(label, h, w)
    keep = ((lab >= 0) & (lab < K))[:, None].double()
    _, S1, _ = class_stats_ref.class_stats(v, label, K)
    xhat = (x.detach() - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    assert torch.allclose(S1.sum(0), (dy * xhat * keep).sum(dim=(0, 2, 3)), rtol=1e-12, atol=1e-12)
    # and without ignored pixels it is autograd's dgamma itself
    _, S1_all, _ = class_stats_ref.class_stats(v, torch.zeros_like(label), 1)
    assert torch.allclose(S1_all[0], bn.weight.grad, rtol=1e-10, atol=1e-12)


def test_scoring_restatement_matches_reference_run():
    """oracle/scoring_ref.py on the workload nets vs the reference's Seg_Model + CriterionDSN + dcfp_pruning
    executed by make_golden.py (2 steps, 2x3x64x128 inputs).  Host convolutions may sum in another order on
    another CPU, so gradients are compared with a tolerance scaled to the layer, EIC likewise."""
    from dcfp_b200.workloads.segnets import build_segnet
    from dcfp_b200.workloads.synthetic import synthetic_batch
    z = np.load(os.path.join(GOLDEN, "scoring_small.npz"))
    K, H, W, steps = int(z["K"]), int(z["H"]), int(z["W"]), int(z["steps"])
    model = build_segnet("deeplabv3", "resnet50", K, seed=0)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        batches = [synthetic_batch([2 * s, 2 * s + 1], K, H, W) for s in range(steps)]
        before = {k: v.clone() for k, v in model.state_dict().items()}
        got, losses = scoring_ref.score(model, batches, r=0.999)
    finally:
        torch.set_num_threads(threads)
    for s in range(steps):
        assert abs(losses[s] - float(z["loss_%d" % s])) <= 1e-5 * abs(float(z["loss_%d" % s]))
    names = [k[5:] for k in z.files if k.startswith("eic::")]
    assert set(names) == set(got["eic"].keys()) and len(names) == 62
    for n in names:
        exp, val = z["eic::" + n], got["eic"][n]
        tol = 1e-4 * np.abs(exp) + 1e-4 * np.abs(exp).mean()
        flipped = (np.abs(val - exp) > tol)
        # a sign-gate flip needs |grad| within rounding of 0; allow none on identical hardware, a handful elsewhere
        assert flipped.mean() <= 0.01, "%s: %d of %d channels differ" % (n, flipped.sum(), flipped.size)
    after = model.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before), "scoring must leave the model untouched"


def _balance_cases():
    import json
    z = np.load(os.path.join(GOLDEN, "balance.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def test_class_balance_restatement_matches_reference_bits():
    """oracle/balance_ref.py vs BaseDataSet.get_label of the reference (datasets/Base.py:73-89): float64, bit-exact."""
    from oracle import balance_ref
    z, cases = _balance_cases()
    for c in cases:
        label = z["label_%d_%d" % (c["case"], c["img"])]
        w, cnt = balance_ref.class_balance_weights(label, c["K"], c["sample_class"], c["balance"], c["beta"])
        exp = z["weight_%d_%d" % (c["case"], c["img"])]
        assert w.dtype == np.float64 and np.array_equal(w.view(np.uint64), exp.view(np.uint64)), c
        assert cnt.sum() == (label != 255).sum()


@pytest.mark.parametrize("relu", [False, True])
def test_factored_batchnorm_equals_autograd(relu):
    """oracle/bn_ref.py (the sum-factored BN (+ReLU) forward/backward a fused BN + K1 kernel pair computes)
    against torch autograd of nn.BatchNorm2d -> nn.ReLU in fp64, including the running-statistics update."""
    from oracle import bn_ref
    torch.manual_seed(3)
    N, C, h, w = 3, 11, 7, 5
    x = (torch.randn(N, C, h, w, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).double().train()
    with torch.no_grad():
        bn.weight.uniform_(-1.5, 1.5)  # negative gammas flip the gate's direction
        bn.bias.normal_()
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    y = bn(x)
    y = torch.relu(y) if relu else y
    dy = torch.randn_like(y)
    y.backward(dy)
    yr, mean, invstd = bn_ref.bn_relu_forward(x.detach(), bn.weight.detach(), bn.bias.detach(), bn.eps, relu=relu)
    assert torch.allclose(yr, y.detach(), rtol=1e-12, atol=1e-12)
    dx, dg, db = bn_ref.bn_relu_backward(x.detach(), dy, bn.weight.detach(), bn.bias.detach(), mean, invstd, relu=relu)
    assert torch.allclose(dx, x.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dg, bn.weight.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(db, bn.bias.grad, rtol=1e-10, atol=1e-12)
    n = N * h * w
    rm, rv = bn_ref.running_stats_update(rm0, rv0, mean, 1.0 / (invstd * invstd) - bn.eps, n, bn.momentum)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-12, atol=1e-12) and torch.allclose(rv, bn.running_var, rtol=1e-10, atol=1e-12)


def test_class_rows_plus_outside_row_equal_autograd_dgamma():
    """The identity the scorer's EIC feed rests on, with the ignore label present: the K class rows of the backward
    functor do NOT add up to bn.weight.grad -- the pixels labelled 255 carry gradient too (they only carry no loss).
    Class rows + the outside row do.  (A product bug of exactly this kind passed every all-valid-label test.)"""
    torch.manual_seed(5)
    N, C, h, w, K = 2, 6, 12, 16, 5
    x = (torch.randn(N, C, h, w, dtype=torch.float64) * 1.5 + 0.3).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).double().train()
    head = torch.nn.Conv2d(C, K, 3, padding=1).double()  # a receptive field: ignored pixels get gradient from neighbours
    label = torch.randint(0, K, (N, 4 * h, 4 * w))
    label[:, 10:30, 5:40] = 255
    y = bn(x)
    logits = F.interpolate(head(torch.relu(y)), size=label.shape[1:], mode="bilinear", align_corners=True)
    y.retain_grad()
    F.cross_entropy(logits, label, ignore_index=255).backward()
    xd = x.detach()
    mean = xd.mean(dim=(0, 2, 3))
    invstd = torch.rsqrt(xd.var(dim=(0, 2, 3), unbiased=False) + bn.eps)
    v = class_stats_ref.functor_bwd(xd, y.grad, invstd, -mean * invstd)
    cnt, S1, S2 = class_stats_ref.class_stats(v, label, K)
    o1, o2 = class_stats_ref.outside_stats(v, label, K)
    g = bn.weight.grad
    assert cnt.sum() < N * h * w and float(o2.sum()) > 0  # some pixels are outside, and they do carry gradient
    assert torch.allclose(S1.sum(0) + o1, g, rtol=1e-9, atol=1e-12)
    assert not torch.allclose(S1.sum(0), g, rtol=1e-3, atol=1e-9), "the class rows alone must NOT reproduce dgamma here"
    assert torch.allclose(S2.sum(0) + o2, (v * v).sum(dim=(0, 2, 3)), rtol=1e-12)
