"""Checks that need the UNMODIFIED reference tree (build container only; skipped on the GPU box):
the workload nets equal the reference's, the product's host logic equals a live reference run, and the
reference's own prune.py runs unmodified against the drop-in `pruners` package."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as gu
from fake_backend import oracle_backend
from oracle import ref_compat

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("arch,K", [("deeplabv3", 19), ("psp", 150), ("deeplabv3p", 171)])
def test_workload_nets_equal_reference(arch, K):
    """Same module names, bit-identical random-init weights and bit-identical CPU outputs."""
    from dcfp_b200.workloads import segnets
    ref = ref_compat.load_reference()
    torch.manual_seed(0)
    m_ref = getattr(ref.networks, arch).Seg_Model(backbone="resnet50", backbone_para=dict(segnets.BACKBONE_PARA), model_para={},
                                                  num_classes=K, align_corner=True, criterion=None, deepsup=True)
    m = segnets.build_segnet(arch, "resnet50", K, seed=0, with_loss=False)
    sd_r, sd = m_ref.state_dict(), m.state_dict()
    assert list(sd_r.keys()) == list(sd.keys())
    assert all(torch.equal(sd_r[k], sd[k]) for k in sd)
    assert m.ignore_prune_layer == m_ref.ignore_prune_layer
    m.eval()
    m_ref.eval()
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        a, b = m_ref(x, deepsup=True), m(x, deepsup=True)
    assert all(torch.equal(i, j) for i, j in zip(a, b))


def test_calibration_loss_equals_reference_criterion():
    import types
    from dcfp_b200.workloads.segnets import CalibrationLoss
    ref = ref_compat.load_reference()
    crit = ref.crit.CriterionDSN(dataset=types.SimpleNamespace(ignore_label=255))
    preds = [torch.randn(2, 19, 16, 16), torch.randn(2, 19, 16, 16)]
    y = torch.randint(0, 19, (2, 16, 16))
    y[0, :3] = 255
    assert torch.equal(crit(preds, y)["loss"], CalibrationLoss()(preds, y)["loss"])


def test_eic_step_live_reference():
    """dcfp_pruning state layout + oracle eic_ref vs the live reference on a model's real gradient layout."""
    from oracle import eic_ref
    ref = ref_compat.load_reference()
    model = gu.build_model("c1")
    tp = ref.dp.dcfp_pruning(model, 0.999)
    from dcfp_b200.pruners import dcfp_pruning
    mine = dcfp_pruning(model, 0.999)
    assert list(tp.state_dict["eic"].keys()) == list(mine.state_dict["eic"].keys())
    assert all(v == 0 and isinstance(v, int) for v in mine.state_dict["eic"].values())
    rng = np.random.RandomState(0)
    state = {n: 0 for n in tp.state_dict["eic"]}
    for step in range(3):
        for n in state:
            m = model.get_submodule(n)
            m.weight.grad = torch.from_numpy((rng.standard_normal(m.weight.numel()) * 1e-4).astype(np.float32))
        tp.step(model)
        for n in state:
            m = model.get_submodule(n)
            state[n] = eic_ref.eic_step(state[n], m.weight.grad.numpy(), m.weight.detach().numpy(), 0.999)
            assert np.array_equal(state[n].view(np.uint32), tp.get_eic()["eic"][n].numpy().view(np.uint32))


def test_prune_model_live_reference_c1():
    ref = ref_compat.load_reference()
    base = gu.build_model("c1", beta_seed=11)
    eic = gu.make_scores(base, "uniform", 21)
    path = "/tmp/_live_score.pth"
    torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, path)
    gp = 0.5 + 0.02 + 0.02 + 0.02
    ma, mb = copy.deepcopy(base), copy.deepcopy(base)
    pr = ref.dp.DCFPPruner(global_percent=gp, layer_keep=0.02, score_file=path)
    sub_r, cfg_r = pr.prune_model(ma, except_start_keys=["conv_deepsup"])
    with oracle_backend():
        from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
        pm = DCFPPruner(global_percent=gp, layer_keep=0.02, score_file=path)
        sub_m, cfg_m = pm.prune_model(mb, except_start_keys=["conv_deepsup"])
        t_m = pm.get_thresh()
    assert [float(t) for t in pr.get_thresh()] == [float(t) for t in t_m]
    assert list(pr.norm_conv_links.items()) == list(pm.norm_conv_links.items())
    assert pr.except_layers == pm.except_layers
    assert list(cfg_r.keys()) == list(cfg_m.keys())
    for k in cfg_r:
        assert list(cfg_r[k].keys()) == list(cfg_m[k].keys()), k
        for kk, a in cfg_r[k].items():
            b = cfg_m[k][kk]
            assert (np.array_equal(a, b) and a.dtype == b.dtype and a.shape == b.shape) if isinstance(a, np.ndarray) else a == b, (k, kk)
    sr, sm = sub_r.state_dict(), sub_m.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    for k in sr:
        if torch.equal(sr[k], sm[k]):
            continue
        assert k.endswith("running_mean") or k in ("last_conv.6.bias", "conv_deepsup.4.bias"), k  # fp32 GEMV of the compensation
        assert torch.allclose(sr[k], sm[k], rtol=1e-4, atol=1e-5), k


def test_reference_prune_py_runs_unmodified_against_dropin(tmp_path):
    """`prune.py` of the reference, byte-for-byte, once with its own `pruners` and once with the drop-in:
    same global_percent trajectory, bit-identical pruned.pth and channel_cfg.pth."""
    model = gu.build_model("c1")
    sd = {k: v for k, v in model.state_dict().items()}
    ckpt, score = str(tmp_path / "model.pth"), str(tmp_path / "score.pth")
    torch.save(sd, ckpt)
    eic = gu.make_scores(model, "uniform", 31)
    torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, score)
    outs = {}
    for which in ("reference", "dropin"):
        save = str(tmp_path / which)
        cmd = [sys.executable, os.path.join(ROOT, "tests", "run_reference_cli.py"), which, "--oracle-backend",
               os.path.join(ref_compat.REF_ROOT, "prune.py"), "--model", "deeplabv3", "--backbone", "resnet50",
               "--backbone-para", '{"os": 8, "mg_unit": [1,2,4], "inplanes": 128, "pretrained": false}',
               "--dataset", "CS", "--prune-ratio", "0.45", "--model-path", ckpt, "--score-path", score, "--save-path", save]
        p = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), timeout=900)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        lines = [l for l in p.stdout.splitlines() if l.startswith(("global_percent", "flops", "Finish"))]
        outs[which] = (lines, torch.load(os.path.join(save, "pruned.pth"), weights_only=False),
                       torch.load(os.path.join(save, "channel_cfg.pth"), weights_only=False))
    assert outs["reference"][0] == outs["dropin"][0] and any(l.startswith("Finish") for l in outs["dropin"][0])
    a, b = outs["reference"][1], outs["dropin"][1]
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
    ca, cb = outs["reference"][2], outs["dropin"][2]
    assert list(ca.keys()) == list(cb.keys())
    for k in ca:
        for kk, v in ca[k].items():
            assert np.array_equal(v, cb[k][kk]) if isinstance(v, np.ndarray) else v == cb[k][kk]
    # the shape-only search (meta-device FLOPs, one K2 call per candidate, ONE gather for the winner) walks the same
    # candidates, prints the same ratios and ends with the same pruned model as the reference's prune.py loop
    from dcfp_b200.pruners.search import prune_to_flops_ratio, search_global_percent
    with oracle_backend():
        gp, cfg_s, trace = search_global_percent(copy.deepcopy(model), score, prune_ratio=0.45)
        sub, cfg_p, gp2 = prune_to_flops_ratio(copy.deepcopy(model), score, prune_ratio=0.45)
    ref_lines = [l for l in outs["reference"][0] if l.startswith("global_percent")]
    assert ["global_percent: {}, flops_ratio: {}".format(g, r) for g, r in trace] == ref_lines
    assert gp == gp2 and "global_percent: {},".format(gp) in ref_lines[-1]
    sd = sub.state_dict()
    assert list(sd.keys()) == list(a.keys()) and all(torch.equal(sd[k], a[k]) for k in a)


@pytest.mark.parametrize("cfg", ["c1", "c3", "c4"])
def test_meta_flops_counter_equals_reference_counter(cfg):
    """dcfp_b200.pruners.flops (forward on the meta device) vs utils/flops_counter.get_model_complexity_info."""
    ref_compat.load_reference()
    from utils.flops_counter import get_model_complexity_info as ref_info
    from dcfp_b200.pruners import flops
    model = gu.build_model(cfg)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exp = ref_info(model, (3, 256, 256), print_per_layer_stat=False, as_strings=False)
        exp_s = ref_info(model, (3, 256, 256), print_per_layer_stat=False)
    assert flops.model_cost(model, (3, 256, 256)) == exp
    assert flops.get_model_complexity_info(model, (3, 256, 256)) == exp_s


def test_random_pruner_live_reference_same_rng_stream():
    """RandomChannelPruner vs the reference's on the same global-RNG seed -- twice, so that the second call of the
    product runs from its topology cache and must still consume the generator like the reference's re-trace does."""
    ref = ref_compat.load_reference()
    from dcfp_b200.pruners.random_pruner import RandomChannelPruner
    base = gu.build_model("c1")
    for seed in (9, 10):
        torch.manual_seed(seed)
        pr = ref.rp.RandomChannelPruner(global_percent=0.7, layer_keep=0.02)
        sa, ca = pr.prune_model(copy.deepcopy(base), except_start_keys=["conv_deepsup"])
        with oracle_backend():
            torch.manual_seed(seed)
            pm = RandomChannelPruner(global_percent=0.7, layer_keep=0.02)
            sb, cb = pm.prune_model(copy.deepcopy(base), except_start_keys=["conv_deepsup"])
        assert all(np.array_equal(ca[k]["out_mask"], cb[k]["out_mask"]) for k in ca)
        assert all(torch.equal(v, sb.state_dict()[k]) for k, v in sa.state_dict().items())


def test_thresholds_and_masks_live_reference_adversarial_scores():
    """`get_thresh` / `gen_channel_mask` (pruners/dcfp_pruner.py:43-92) of the LIVE reference against the product's single
    select call (oracle backend on CPU) on score sets built to hit the rules the golden sweeps only touch by accident:
    heavily tied (quantised) scores, all-equal and all-zero scores, a threshold sitting on a tie, the min-keep fallback
    on every layer, layer_keep = 0 (-> 1 channel) and large layer_keep, global_percent at both ends of its range.
    Both pruners are prepared once (the 10 s trace) and re-used: thresholds and masks are pure functions of
    (scores, global_percent, layer_keep) afterwards."""
    ref = ref_compat.load_reference()
    base = gu.build_model("c1")
    names = [n for n, m in base.named_modules() if isinstance(m, torch.nn.BatchNorm2d) and n not in base.ignore_prune_layer]
    sizes = {n: base.get_submodule(n).weight.numel() for n in names}
    path = "/tmp/_live_adv_score.pth"
    torch.save({"eic": {n: torch.zeros(c) for n, c in sizes.items()}}, path)

    def prepared(cls):
        pr = cls(global_percent=0.5, layer_keep=0.02, score_file=path)
        m = copy.deepcopy(base)
        pr.end_nodes = getattr(m, "end_nodes", [])
        pr.prepare_from_supernet(m)
        pr.except_start_keys = pr.except_start_keys + m.ignore_prune_layer + ["conv_deepsup"]
        pr.get_except_layers(m)
        return pr

    from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
    pr, pm = prepared(ref.dp.DCFPPruner), prepared(DCFPPruner)
    assert list(pr.norm_conv_links.items()) == list(pm.norm_conv_links.items()) and pr.except_layers == pm.except_layers
    rng = np.random.RandomState(5)

    def scores(kind):
        out = {}
        for n, c in sizes.items():
            if kind == "quantised":      # 8 distinct values: every threshold sits inside a run of ties
                s = rng.randint(0, 8, c).astype(np.float32) * 0.125
            elif kind == "all_equal":
                s = np.full(c, 0.25, np.float32)
            elif kind == "all_zero":
                s = np.zeros(c, np.float32)
            elif kind == "one_hot":      # a single non-zero channel per layer: masks come from the min-keep fallback almost everywhere
                s = np.zeros(c, np.float32)
                s[rng.randint(c)] = rng.rand() + 0.5
            elif kind == "denormal":     # scores around the smallest fp32 normals and below (an EIC that was gated off for 10^4 steps)
                s = (rng.rand(c) * 4e-38).astype(np.float32)
                s[rng.rand(c) < 0.5] = 0.0
            else:                        # "signed": the EIC is never negative, the pruner must not assume so
                s = rng.standard_normal(c).astype(np.float32)
            out[n] = torch.from_numpy(s)
        return out

    cases = [(k, gp, lk) for k in ("quantised", "all_equal", "all_zero", "one_hot", "denormal", "signed")
             for gp, lk in ((0.0, 0.02), (0.5, 0.0), (0.62, 0.02), (0.9, 0.5), (0.999, 0.02))]
    tie_breaks = 0
    with oracle_backend():
        for kind, gp, lk in cases:
            eic = scores(kind)
            for p in (pr, pm):
                p.eic, p.global_percent, p.layer_keep = eic, gp, lk
            t_r = [float(t) for t in pr.get_thresh()]
            t_m = [float(t) for t in pm.get_thresh()]
            assert np.array_equal(np.float32(t_r).view(np.uint32), np.float32(t_m).view(np.uint32)), (kind, gp, lk, t_r, t_m)
            pr.gen_channel_mask()
            pm.gen_channel_mask()
            for bn, conv in pr.norm_conv_links.items():
                if conv in pr.except_layers:
                    continue
                a, b = pr.name2module[conv].out_mask, pm.name2module[conv].out_mask
                assert a.shape == b.shape and a.dtype == b.dtype, (kind, gp, lk, bn)
                if torch.equal(a, b):
                    continue
                # The only licensed difference: the min-keep fallback (:82-84) takes `torch.sort(score, descending=True)[:k]`,
                # an UNSTABLE sort -- among channels whose score equals the k-th largest, which ones it returns is an
                # implementation detail of torch's CPU sort (n > 16).  The product keeps the lowest indices among them
                # (= a stable sort).  Same number of channels, same kept scores, differences only inside that tie.
                tie_breaks += 1
                sc, fa, fb = eic[bn], a.reshape(-1) == 1, b.reshape(-1) == 1
                assert int(fa.sum()) == int(fb.sum()) == max(int(sc.numel() * lk), 1), (kind, gp, lk, bn)
                assert torch.equal(sc[fa].sort().values, sc[fb].sort().values), (kind, gp, lk, bn)
                cut = sc[fa].min()
                assert bool((sc[fa != fb] == cut).all()) and int((sc == cut).sum()) > int((sc[fa] == cut).sum()), (kind, gp, lk, bn)
    assert tie_breaks > 0  # the quantised / all-equal sets do reach the rule
    with pytest.raises(IndexError):  # global_percent = 1.0 indexes one past the sorted scores: both raise
        pr.global_percent = 1.0
        pr.get_thresh()
    with oracle_backend(), pytest.raises(IndexError):
        pm.global_percent = 1.0
        pm.get_thresh()
