"""DCFPPruner.prune_model through the REAL CUDA kernels (K2 thresholds/masks, K3 grouped gather, bias-compensation
GEMV) against the golden outputs of the unmodified reference (tests/golden/prune_*.npz): thresholds, kept-channel
index sets and every pruned tensor BIT-EXACT on all four BASELINE models."""
import copy

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_prune_model_bit_exact_vs_reference_golden(native, cfg, device):
    """device: where the model lives when handed to prune_model (prune.py keeps it on the CPU; a GPU-resident model
    skips the staging copies)."""
    from dcfp_b200 import ops
    z, meta = gu.load_fixture(cfg)
    base = gu.build_model(cfg)
    for ci, case in enumerate(meta["cases"]):
        eic = gu.make_scores(base, case["kind"], case["seed"])
        assert gu.scores_digest(eic) == case["scores_sha256"]
        model = copy.deepcopy(base).to(device)
        n0 = ops.launch_count()
        pruner, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
        assert ops.launch_count() - n0 >= 3, "the CUDA kernels must be what runs"
        assert next(sub.parameters()).device.type == device
        gu.check_case(z, meta, ci, pruner, sub, ccfg)


def test_bias_compensation_vs_reference_golden(native):
    z, meta = gu.load_fixture("c1_beta")
    model = gu.build_model("c1", beta_seed=3)
    case = meta["cases"][0]
    eic = gu.make_scores(model, case["kind"], case["seed"])
    pruner, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
    sd = gu.check_case(z, meta, 0, pruner, sub, ccfg, check_topology=False, skip=meta["running_means"])
    for k in meta["running_means"]:
        exp, got = z["rm::" + k], sd[k].cpu().numpy()
        # fp32 reduce-GEMV, summation order differs from MKL's: 1e-4 relative to the offsets' scale
        assert np.allclose(got, exp, rtol=1e-4, atol=1e-4 * max(np.abs(exp).max(), 1e-3)), k


def test_pruned_model_runs(native):
    """The sliced network is consistent (every in/out channel count fits its neighbours) and runs."""
    z, meta = gu.load_fixture("c1")
    case = meta["cases"][0]
    base = gu.build_model("c1").eval()
    eic = gu.make_scores(base, case["kind"], case["seed"])
    model = copy.deepcopy(base)
    _, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
    sub = sub.eval()
    x = torch.randn(1, 3, 64, 128)
    with torch.no_grad():
        out = sub(x, deepsup=True)
    assert out[0].shape == (1, 19, 64, 128) and torch.isfinite(out[0]).all()
    kept = sum(c["out_channels"] for c in ccfg.values())
    raw = sum(c["raw_out_channels"] for c in ccfg.values())
    assert kept < raw


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_global_percent_sweep_bit_exact(native, cfg):
    """K2 (radix-select thresholds, strict-> masks, min-keep fallback) for all 25 global_percent values prune.py can
    visit, eic-like scores with 40 % exact zeros: thresholds and masks bit-exact vs the unmodified reference."""
    gu.check_percent_sweep(cfg)
