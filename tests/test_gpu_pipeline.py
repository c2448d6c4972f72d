"""End to end on the GPU: calibration scoring (K1 + K2a) -> score.pth -> FLOPs-ratio search -> prune_model (K2b + K3), then
the SAME score file through the oracle backend (CPU restatements of thresholds / masks / gather pinned by the reference's
golden outputs): the two pruned models must be identical -- kept-channel index sets and every tensor bit-exact (SURVEY.md
section 7.3: parity on shared bits; only K2/K3 differ between the two runs)."""
import copy

import numpy as np
import pytest
import torch

from fake_backend import oracle_backend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,classes", [("deeplabv3", 19), ("psp", 150)])
def test_score_then_prune_identical_to_oracle_backend(native, tmp_path, arch, classes):
    from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
    from dcfp_b200.pruners.search import prune_to_flops_ratio
    from dcfp_b200.scorer import score_calibration_set
    from dcfp_b200.workloads.segnets import build_segnet
    from dcfp_b200.workloads.synthetic import synthetic_batch

    base = build_segnet(arch, "resnet50", classes, seed=0)
    x, y = synthetic_batch(list(range(8)), classes, 128, 256)
    gpu_model = copy.deepcopy(base).to("cuda").to(memory_format=torch.channels_last)
    out = score_calibration_set(gpu_model, x, y, classes, micro_batch=2, r=0.999, seed=0)
    score = str(tmp_path / "score.pth")
    torch.save({"eic": out["eic"]}, score)  # the reference's score.pth layout (pruners/dcfp_pruner.py:25-26)
    flat = np.concatenate([v.numpy() for v in out["eic"].values()])
    assert np.isfinite(flat).all() and (flat >= 0).all() and (flat > 0).mean() > 0.5

    # GPU path: search + one prune_model (K2b thresholds/masks, K3 gather)
    m_gpu = copy.deepcopy(base)
    m_gpu.criterion = None
    sub_gpu, cfg_gpu, gp = prune_to_flops_ratio(m_gpu, score, prune_ratio=0.5)
    # oracle backend on the very same score bits
    m_cpu = copy.deepcopy(base)
    m_cpu.criterion = None
    with oracle_backend():
        pruner = DCFPPruner(global_percent=gp, layer_keep=0.02, score_file=score)
        sub_cpu, cfg_cpu = pruner.prune_model(m_cpu, except_start_keys=["conv_deepsup"])
    assert list(cfg_gpu.keys()) == list(cfg_cpu.keys())
    for k in cfg_gpu:
        for kk, v in cfg_gpu[k].items():
            w = cfg_cpu[k][kk]
            assert np.array_equal(v, w) if isinstance(v, np.ndarray) else v == w, (k, kk)
    sa, sb = sub_gpu.state_dict(), sub_cpu.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k].cpu(), sb[k].cpu()) for k in sa)
    # a second scoring stack on the same images -- torch/cuDNN BatchNorm + ReLU with the hook-fed K1 instead of the fused
    # BN kernels.  At random init the BN-gamma gradients are noise-like (tests/test_gpu_fused_scorer.py: plain fp32 torch is
    # 3-10 % away from an fp64 arbiter per tensor) and the EIC's sign gate turns a sign flip of a near-zero gradient into
    # an O(1) change of that channel's score, so two stacks agree on MOST channels only.  The masks are therefore compared
    # WITH the threshold margin: how close the nearest score is to each threshold, how many keep bits
    # differ, and how far from the threshold those channels sit.
    from dcfp_b200.pruners.margin import compare_masks, format_margins
    out_b = score_calibration_set(copy.deepcopy(base).to("cuda").to(memory_format=torch.channels_last), x, y, classes,
                                  micro_batch=2, r=0.999, seed=0, fused=False)
    names = list(out["eic"].keys())
    sizes = [out["eic"][n].numel() for n in names]
    offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    grp = [0 if n.startswith("backbone") else 1 for n in names]
    cmp = compare_masks(flat, np.concatenate([out_b["eic"][n].numpy() for n in names]), offs, grp, gp)
    print("fused vs unfused stack @ global_percent %.2f: %s | %s; %d / %d keep bits differ, all within %.3g of their "
          "threshold; score discrepancy near the thresholds %.3g" % (gp, format_margins(cmp["margins_a"]),
          format_margins(cmp["margins_b"]), cmp["flipped"], cmp["n"], cmp["flip_band"], cmp["disc_near"]))
    assert cmp["flipped"] <= 0.25 * cmp["n"], "the two stacks disagree on more keep bits than gradient noise explains"
    assert all(v["margin"] >= 0 for v in cmp["margins_a"].values())
    kept = sum(c["out_channels"] for c in cfg_gpu.values())
    raw = sum(c["raw_out_channels"] for c in cfg_gpu.values())
    assert kept < raw
    sub_gpu.eval()
    with torch.no_grad():
        pred = sub_gpu(torch.randn(1, 3, 64, 64), deepsup=True)
    assert pred[0].shape == (1, classes, 64, 64) and torch.isfinite(pred[0]).all()
