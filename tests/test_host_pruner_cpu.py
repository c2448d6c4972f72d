"""HOST logic of the product's pruner mirror (graph analysis, mask propagation, bias / gather
bookkeeping) against the golden outputs of the unmodified reference.  The CUDA kernels are replaced by
the CPU oracle through tests/fake_backend.py (test-only): what is pinned here is (a) the host logic and
(b) the oracle's mask / gather restatements.  The same fixtures are checked against the real CUDA
kernels by tests/test_gpu_prune_golden.py."""
import numpy as np
import pytest
import torch

import golden_util as gu
from fake_backend import oracle_backend


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_prune_model_matches_reference_golden(cfg):
    z, meta = gu.load_fixture(cfg)
    base = gu.build_model(cfg)
    base_sd = {k: v.clone() for k, v in base.state_dict().items()}
    n_cases = len(meta["cases"]) if cfg == "c1" else 1  # R101 traces are slow on CPU: one case each here, all on the GPU
    import copy
    for ci in range(n_cases):
        case = meta["cases"][ci]
        eic = gu.make_scores(base, case["kind"], case["seed"])
        assert gu.scores_digest(eic) == case["scores_sha256"], "score generator drifted from the fixture"
        model = copy.deepcopy(base)
        with oracle_backend():
            pruner, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
            assert sub is model  # prune_model mutates and returns its argument (reference :967-990)
            gu.check_case(z, meta, ci, pruner, sub, ccfg)
    assert all(torch.equal(base_sd[k], v) for k, v in base.state_dict().items())


def test_bias_compensation_matches_reference_golden():
    """non-zero BN beta: consumers of pruned channels get running_mean -= W.sum((2,3)) @ relu(beta) (reference :873-905)."""
    z, meta = gu.load_fixture("c1_beta")
    model = gu.build_model("c1", beta_seed=3)
    before = {k: v.clone() for k, v in model.state_dict().items() if k.endswith("running_mean")}
    case = meta["cases"][0]
    eic = gu.make_scores(model, case["kind"], case["seed"])
    assert gu.scores_digest(eic) == case["scores_sha256"]
    with oracle_backend():
        pruner, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
        sd = gu.check_case(z, meta, 0, pruner, sub, ccfg, check_topology=False, skip=meta["running_means"])
    moved = 0
    for k in meta["running_means"]:
        exp = z["rm::" + k]
        got = sd[k].numpy()
        assert got.shape == exp.shape
        # fp32 GEMV (MKL in the reference, fp64 in the oracle): tolerance relative to the offsets' scale
        assert np.allclose(got, exp, rtol=1e-4, atol=1e-4 * max(np.abs(exp).max(), 1e-3)), k
        moved += int(np.abs(exp).max() > 0)
    assert moved > 30, "fixture should exercise the compensation on many layers"


def test_init_pruned_model_roundtrip():
    """channel_cfg -> init_pruned_model on a FRESH model -> load_state_dict(pruned) (prune.py:100-110)."""
    from dcfp_b200.pruners import init_pruned_model
    z, meta = gu.load_fixture("c1")
    case = meta["cases"][0]
    model = gu.build_model("c1")
    eic = gu.make_scores(model, case["kind"], case["seed"])
    with oracle_backend():
        _, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
    fresh = gu.build_model("c1")
    init_pruned_model(fresh, ccfg)
    missing = fresh.load_state_dict(sub.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    fresh.eval()
    sub.eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        a, b = fresh(x, deepsup=True), sub(x, deepsup=True)
    assert all(torch.equal(i, j) for i, j in zip(a, b))


def test_random_pruner_and_min_keep():
    from dcfp_b200.pruners.random_pruner import RandomChannelPruner
    model = gu.build_model("c1")
    torch.manual_seed(5)
    with oracle_backend():
        pruner = RandomChannelPruner(global_percent=0.999, layer_keep=0.02)
        sub, ccfg = pruner.prune_model(model, except_start_keys=["conv_deepsup"])
    for name, c in ccfg.items():
        assert c["out_channels"] >= 1
    c = ccfg["backbone.layer1.0.conv1"]
    assert c["out_channels"] >= max(int(c["raw_out_channels"] * 0.02), 1)
    sub.eval()
    with torch.no_grad():
        out = sub(torch.randn(1, 3, 64, 64), deepsup=True)
    assert out[0].shape == (1, 19, 64, 64)


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_global_percent_sweep_matches_reference_golden(cfg):
    """oracle backend: pins oracle/mask_ref.py and the host bookkeeping of DCFPPruner._select for every percent."""
    with oracle_backend():
        gu.check_percent_sweep(cfg)


def test_meta_flops_counter_golden_numbers():
    """GFLOPs the UNMODIFIED reference counter reported for the four BASELINE models at 3x512x512 (prune.py:78; measured
    in the build container) -- reproduced from shapes alone on the meta device."""
    from dcfp_b200.pruners import flops
    from dcfp_b200.workloads.segnets import CONFIGS, build_segnet
    expected = {"c1": ("177.29 GFLOPs", "41.27 M"), "c2": ("255.19 GFLOPs", "60.26 M"), "c3": ("262.79 GFLOPs", "65.77 M"),
                "c4": ("279.66 GFLOPs", "60.42 M")}
    for cfg, exp in expected.items():
        c = CONFIGS[cfg]
        model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0, with_loss=False, deepsup=False)
        assert flops.get_model_complexity_info(model, (3, 512, 512)) == exp, cfg


def test_search_global_percent_oracle_backend():
    """The FLOPs-ratio search stops at the first candidate under the target and its channel_cfg is the one prune_model
    produces at that percent."""
    from dcfp_b200.pruners.search import search_global_percent
    model = gu.build_model("c1")
    eic = gu.make_scores(model, "uniform", 31)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".pth") as f:
        torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, f.name)
        with oracle_backend():
            gp, cfg_s, trace = search_global_percent(model, f.name, prune_ratio=0.62)
            _, _, cfg_p = gu.run_product_prune(__import__("copy").deepcopy(model), eic, gp, 0.02)
    ratios = [r for _, r in trace]
    assert all(r > 0.38 for r in ratios[:-1]) and ratios[-1] <= 0.38 and len(trace) >= 3
    assert [g for g, _ in trace][:3] == [0.5, 0.52, 0.54]
    assert ratios == sorted(ratios, reverse=True)
    for k in cfg_p:
        assert cfg_p[k]["out_channels"] == cfg_s[k]["out_channels"] and cfg_p[k].get("in_channels") == cfg_s[k].get("in_channels")
        assert np.array_equal(cfg_p[k]["out_mask"], cfg_s[k]["out_mask"])
