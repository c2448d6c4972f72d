"""The scorer with the fused BatchNorm(+ReLU) kernels: a channels_last model is scored with its BN layers
running on dcfp's own forward / backward, the class rows coming out of the BN backward itself.  Checked here:
  * the fused pass is the same FUNCTION as torch's BN + ReLU: losses, every parameter gradient and the EIC scores agree
    with the unfused (cuDNN BN + hook + deferred K1) pass of the same model on the same micro-batches,
  * sum over the class rows == the bn.weight.grad autograd hands out (now both produced by one kernel: exact to fp32),
  * the model is left untouched: modules unpatched, running statistics and counters restored,
  * which layers fuse with their ReLU (bn1 / bn2 of a bottleneck) and which with the residual add + ReLU behind them (bn3:
    the lazy.PendingBN protocol, tests/test_lazy_bn_cpu.py), and that the latter changes nothing but the kernel count.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
K, H, W = 19, 128, 256


def _model(arch="deeplabv3", classes=K, seed=0):
    from dcfp_b200.workloads.segnets import build_segnet
    m = build_segnet(arch, "resnet50", classes, seed=seed).to(DEV).to(memory_format=torch.channels_last)
    for mod in m.modules():  # CPU/CUDA-independent comparison: no dropout noise between the two passes
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
    return m


def _batch(idx, classes=K):
    from dcfp_b200.workloads.synthetic import synthetic_batch
    x, y = synthetic_batch(idx, classes, H, W)
    return x.to(DEV).contiguous(memory_format=torch.channels_last), y.to(DEV)


def _run(model, classes, fused, steps=3, keep_grads=False, fuse_residual=True):
    from dcfp_b200.scorer import CalibrationRun
    run = CalibrationRun(model, classes, r=0.999, seed=5, fused=fused, fuse_residual=fuse_residual)
    losses, grads = [], None
    for s in range(steps):
        x, y = _batch([2 * s, 2 * s + 1], classes)
        losses.append(float(run.step(x, y, mb_index=s)))
        if keep_grads and s == steps - 1:
            grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    sc = run.scorer
    eic = sc.eic.cpu().numpy().copy()
    totals = sc.totals.cpu().clone()
    info = dict(fused_calls=sc.fused_layer_calls, relu_after=dict(sc._relu_after), add_relu_after=dict(sc._add_relu_after),
                tail_calls=sc.fused_tail_calls)
    run.close()
    return losses, eic, totals, grads, info


def _arbiter_grads(model, classes, step):
    """fp64 ARBITER: the same network in double precision under plain autograd (torch's own BN / ReLU) on the same
    micro-batch.  At random init the gradients below the first BN backward are hypersensitive (heavy cancellation): plain
    fp32 torch is 3-10 % away from this arbiter per parameter tensor (scripts/check_fused_vs_unfused.py), so 'fused equals
    unfused' can only be judged by the distance of each to the arbiter."""
    import copy
    arb = copy.deepcopy(model).double().train()
    x, y = _batch([2 * step, 2 * step + 1], classes)
    out = arb(x.double(), y.long(), deepsup=True)
    (out["loss"] if isinstance(out, dict) else out).backward()
    return float(out["loss"].detach()), {n: p.grad.detach().clone() for n, p in arb.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("arch,classes", [("deeplabv3", 19), ("psp", 150), ("deeplabv3p", 171)])
def test_fused_pass_equals_unfused_pass(native, arch, classes):
    model = _model(arch, classes)
    before = {n: b.detach().clone() for n, b in model.named_buffers()}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # IEEE fp32 convolutions: the comparison is about BN, not about tf32 noise
    try:
        la, ga = _arbiter_grads(model, classes, step=2)
        l0, e0, t0, g0, i0 = _run(model, classes, fused=False, keep_grads=True)
        l1, e1, t1, g1, i1 = _run(model, classes, fused=True, keep_grads=True)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert i0["fused_calls"] == 0 and i1["fused_calls"] > 100, i1["fused_calls"]
    assert np.allclose(l0, l1, rtol=2e-5) and abs(l1[-1] - la) < 2e-5 * abs(la), (l0, l1, la)
    # every parameter gradient of the last step against the fp64 arbiter: the fused pass must be as close to it as
    # torch's own fp32 BN / ReLU is
    ratios = []
    for n in ga:
        ref = ga[n].abs().max() + 1e-30
        ef = float((g1[n].double() - ga[n]).abs().max() / ref)
        eu = float((g0[n].double() - ga[n]).abs().max() / ref)
        ratios.append(ef / (eu + 1e-4))
        assert ef <= 10.0 * eu + 2e-3, "%s: fused %.3g vs unfused %.3g away from the fp64 arbiter" % (n, ef, eu)
    # per tensor the two distances are noise of the same size (either can be the larger one): judge the distribution
    assert np.median(ratios) < 1.3 and np.quantile(ratios, 0.9) < 2.5, (np.median(ratios), np.quantile(ratios, 0.9), max(ratios))
    # EIC scores of the two passes: noise-like gradients + a sign gate, so distribution-level (as the oracle test)
    rel = np.abs(e1 - e0) / (np.abs(e0) + 0.1 * np.abs(e0).mean())
    q50, q90 = np.quantile(rel, [0.5, 0.9])
    assert q50 < 5e-2 and q90 < 0.25 and (rel <= 0.5).mean() > 0.97, (q50, q90)
    # the model is back: no instance-level forwards, buffers restored bit for bit
    assert not any("forward" in mod.__dict__ for mod in model.modules())
    for n, b in model.named_buffers():
        assert torch.equal(b, before[n]), n
    print("fused / unfused distance to the fp64 arbiter: median ratio %.3f, max %.3f; score q50 %.3g q90 %.3g" %
          (np.median(ratios), max(ratios), q50, q90))


def test_which_layers_fuse_with_their_relu(native):
    model = _model()
    _, _, _, _, info = _run(model, K, fused=True, steps=2)
    ra = info["relu_after"]
    assert ra.get("backbone.layer1.0.bn1") and ra.get("backbone.layer1.0.bn2") and ra.get("backbone.bn1")
    assert "backbone.layer1.0.bn3" not in ra and "backbone.layer1.0.downsample.1" not in ra  # residual add comes first
    assert ra.get("aspp.aspp1.bn") and ra.get("last_conv.1")
    assert ra.get("aspp.bn1")  # in ignore_prune_layer: not scored, but run by the same kernels (class rows to a scratch arena)
    # bn3 -> (+ shortcut) -> ReLU: learned in step 1, one fused call per bottleneck in step 2 (ResNet-50: 16 blocks; the last
    # one, layer4.2.bn3, is in ignore_prune_layer -- unscored, same kernels)
    ar = info["add_relu_after"]
    assert ar.get("backbone.layer1.0.bn3") and ar.get("backbone.layer3.5.bn3") and ar.get("backbone.layer4.2.bn3")
    assert "backbone.layer1.0.downsample.1" not in ar and "backbone.layer1.0.bn1" not in ar
    assert len(ar) == 16 and info["tail_calls"] == 16, (len(ar), info["tail_calls"])
    _, _, _, _, off = _run(model, K, fused=True, steps=2, fuse_residual=False)
    assert off["add_relu_after"] == {} and off["tail_calls"] == 0


def test_residual_tail_fusion_changes_no_result(native):
    """Same model, same micro-batches, IEEE fp32 convolutions: with and without the fused bn3 + shortcut + ReLU tail the
    losses agree to fp32 round-off, and gradients / pass totals differ by no more than two runs of the SAME program do
    (statistics and sums come from atomics, whose order changes from run to run; a random-init net amplifies that)."""
    model = _model(seed=3)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        l0, e0, t0, g0, i0 = _run(model, K, fused=True, steps=3, keep_grads=True, fuse_residual=False)
        l2, e2, t2, g2, i2 = _run(model, K, fused=True, steps=3, keep_grads=True, fuse_residual=False)
        l1, e1, t1, g1, i1 = _run(model, K, fused=True, steps=3, keep_grads=True, fuse_residual=True)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert i1["tail_calls"] == 32 and i0["tail_calls"] == 0 and i0["fused_calls"] == i1["fused_calls"]
    assert np.allclose(l0, l1, rtol=2e-6), (l0, l1)

    def dist(ga, gb):
        return np.array([float((ga[n] - gb[n]).abs().max() / (gb[n].abs().max() + 1e-30)) for n in gb])

    d_same, d_tail = dist(g2, g0), dist(g1, g0)
    assert np.median(d_tail) <= 3.0 * np.median(d_same) + 1e-5 and d_tail.max() <= 3.0 * d_same.max() + 1e-3, \
        (np.median(d_tail), np.median(d_same), d_tail.max(), d_same.max())
    tot_same = np.abs(t2.numpy() - t0.numpy()).sum() / np.abs(t0.numpy()).sum()
    tot_tail = np.abs(t1.numpy() - t0.numpy()).sum() / np.abs(t0.numpy()).sum()
    assert tot_tail <= 3.0 * tot_same + 1e-5, (tot_tail, tot_same)
    print("gradient distance to a reference run: same program median %.3g max %.3g | fused tail median %.3g max %.3g; totals %.3g | %.3g"
          % (np.median(d_same), d_same.max(), np.median(d_tail), d_tail.max(), tot_same, tot_tail))


def test_row_sums_are_the_gradient_autograd_hands_out(native):
    from dcfp_b200.scorer import CalibrationRun
    model = _model(seed=2)
    run = CalibrationRun(model, K, r=0.999, seed=1, fused=True)
    for s in range(2):
        x, y = _batch([2 * s, 2 * s + 1])
        run.step(x, y, mb_index=s)
    sc = run.scorer
    # the step arena was folded: recover this step's dgamma from the EIC input instead -- run one more step by hand
    x, y = _batch([4, 5])
    sc.set_labels(y)
    model.zero_grad(set_to_none=True)
    out = model(x, y.long(), deepsup=True)  # noqa
    (out["loss"] if isinstance(out, dict) else out).backward()
    sc.flush()
    both = sc.step_arena[0] + sc.step_arena32[0].double()  # hook-fed fp64 rows + the fused layers' fp32 rows
    rows = both.sum(0).cpu().numpy()
    grads = torch.cat([m.weight.grad.reshape(-1) for _, m in sc.layers]).cpu().numpy()
    mass = both.abs().sum(0).cpu().numpy()
    # fused layers: rows and gradient come out of ONE kernel (fp64 totals rounded once).  The few maps the fused path does
    # not take (1x1 pooled maps) keep torch's BN: there the gradient is cuDNN's own fp32 reduction (SURVEY app. C form)
    fused = np.concatenate([np.full(m.weight.numel(), "forward" in m.__dict__ and m.weight.numel() % 4 == 0) for _, m in sc.layers])
    big = np.concatenate([np.full(m.weight.numel(), n != "aspp.global_avg_pool.2") for n, m in sc.layers])
    err = np.abs(rows - grads)
    assert (err[fused & big] <= 1e-6 * mass[fused & big] + 1e-12).all()
    assert (err[~(fused & big)] <= 2e-4 * np.abs(grads[~(fused & big)]) + 2e-4 * np.abs(grads).mean()).all()
    assert sum("forward" in m.__dict__ for _, m in sc.layers) >= 60
    run.close()


def test_graph_replay_equals_eager_launches(native):
    """The third step of a shape is captured into a CUDA graph and replayed from then on: same kernels, same results as
    launching them eagerly; a cached run (score_calibration_set) replays on its first step."""
    from dcfp_b200 import scorer
    from dcfp_b200.scorer import CalibrationRun, score_calibration_set
    model = _model(seed=3)
    res = {}
    for graph in (False, True):
        run = CalibrationRun(model, K, r=0.999, seed=2, fused=True, graph=graph)
        losses = []
        for s in range(5):
            x, y = _batch([2 * (s % 2), 2 * (s % 2) + 1])
            losses.append(float(run.step(x, y, mb_index=s)))
        res[graph] = (losses, run.scorer.eic.cpu().numpy().copy(), run.scorer.totals.cpu().clone(), run.graph_replays)
        run.close()
    assert res[False][3] == 0 and res[True][3] == 3
    assert np.allclose(res[False][0], res[True][0], rtol=1e-6)
    rel = np.abs(res[True][1] - res[False][1]) / (np.abs(res[False][1]) + 0.1 * np.abs(res[False][1]).mean())
    assert np.quantile(rel, 0.5) < 1e-3 and (rel < 0.5).mean() > 0.98, np.quantile(rel, [0.5, 0.9, 0.99])  # cuDNN atomics, noise-like gradients
    # the run cache: second call on the same model skips capture and gives the same scores as the first
    from dcfp_b200.workloads.synthetic import synthetic_batch
    xs, ys = synthetic_batch(list(range(8)), K, H, W)
    a = score_calibration_set(model, xs, ys, K, seed=0)
    assert id(model) in scorer._RUN_CACHE
    b = score_calibration_set(model, xs, ys, K, seed=0)
    ea, eb = np.concatenate([v.numpy() for v in a["eic"].values()]), np.concatenate([v.numpy() for v in b["eic"].values()])
    rel = np.abs(ea - eb) / (np.abs(ea) + 0.1 * np.abs(ea).mean())
    assert np.quantile(rel, 0.5) < 1e-3 and (rel < 0.5).mean() > 0.98
    scorer.release_cached_runs()
    assert not scorer._RUN_CACHE and not any("forward" in m.__dict__ for m in model.modules())


def test_eval_mode_and_no_grad_fall_back_to_torch(native):
    """Outside a scoring step (no labels set / eval mode / no_grad) a patched BN is torch's BN."""
    from dcfp_b200.scorer import ClassStatsScorer
    model = _model()
    sc = ClassStatsScorer(model, K, fused=True).attach()
    x, _ = _batch([0, 1])
    model.eval()
    with torch.no_grad():
        a = model(x)[0]
    sc.detach()
    with torch.no_grad():
        b = model(x)[0]
    assert torch.equal(a, b) and sc.fused_layer_calls == 0
