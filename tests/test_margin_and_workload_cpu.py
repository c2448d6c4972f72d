"""Host-side helpers added in round 2: threshold margins, the label-fragmentation presets of the synthetic
workload, and bench.py's `config` dict being the same object from both arms."""
import importlib.util
import os
import types

import numpy as np
import torch

from dcfp_b200.pruners.margin import compare_masks, format_margins, threshold_margins
from dcfp_b200.workloads.synthetic import FRAGMENTATION, label_run_stats, synthetic_batch, synthetic_labels
from oracle import mask_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_threshold_margins_follow_the_reference_threshold_rule():
    rng = np.random.RandomState(3)
    sizes = [64, 128, 256, 48]
    groups = [0, 0, 1, 1]
    layers = [rng.rand(c).astype(np.float32) for c in sizes]
    score = np.concatenate(layers)
    off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    for gp in (0.5, 0.62, 0.9):
        m = threshold_margins(score, off, groups, gp)
        t_ref = mask_ref.thresholds(layers, groups, gp)  # the oracle's restatement of dcfp_pruner.py:43-66
        for g in (0, 1):
            assert np.float32(m[g]["thresh"]) == np.float32(t_ref[g])
            vals = np.sort(np.concatenate([l for l, gg in zip(layers, groups) if gg == g]))
            gap = min(m[g]["thresh"] - vals[vals < m[g]["thresh"]].max(), vals[vals > m[g]["thresh"]].min() - m[g]["thresh"])
            assert abs(m[g]["margin"] - gap / m[g]["thresh"]) < 1e-6
    assert "group 0" in format_margins(m)


def test_compare_masks_counts_flips_inside_the_discrepancy_band():
    rng = np.random.RandomState(0)
    a = rng.rand(4000).astype(np.float32)
    off, grp = [0, 1500, 4000], [0, 1]
    same = compare_masks(a, a.copy(), off, grp, 0.5)
    assert same["flipped"] == 0 and same["flip_band"] == 0.0
    b = (a * (1 + 2e-3 * rng.randn(4000))).astype(np.float32)
    c = compare_masks(a, b, off, grp, 0.5)
    assert 0 < c["flipped"] < 40 and c["flip_band"] <= 4 * c["disc_near"]
    # half of the scores exactly zero (one EIC step): the threshold is 0, the margin is absolute
    z = a.copy()
    z[rng.rand(4000) < 0.6] = 0.0
    m = threshold_margins(z, off, grp, 0.5)
    assert m[0]["thresh"] == 0.0 and not m[0]["relative"] and m[0]["margin"] == z[:1500][z[:1500] > 0].min()


def test_fragmentation_presets_are_deterministic_and_ordered():
    base = synthetic_labels(5, 19, 128, 256)
    assert torch.equal(base, synthetic_labels(5, 19, 128, 256, fragmentation="coarse"))  # the default IS the coarse preset
    runs = {}
    for name in FRAGMENTATION:
        a = synthetic_labels(7, 19, 256, 512, fragmentation=name)
        assert torch.equal(a, synthetic_labels(7, 19, 256, 512, fragmentation=name))
        assert a.dtype == torch.uint8 and 0.01 < float((a == 255).float().mean()) < 0.1
        assert int(a[a != 255].max()) < 19
        runs[name] = label_run_stats(a[None])[0]
    assert runs["coarse"] > runs["street"] > runs["fine"] > 1.0
    x, y = synthetic_batch([0, 1], 150, 64, 64, fragmentation="street")
    assert x.shape == (2, 3, 64, 64) and y.shape == (2, 64, 64)


def test_both_bench_arms_print_the_same_config():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    args = types.SimpleNamespace(micro_batch=2, labels="street")
    c = bench.workload("c2")
    cfg = bench.shared_config(c, args)
    assert cfg == bench.shared_config(bench.workload("c2"), args) and "configs[1]" in cfg["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": shared_config(c, args)') == 2  # the reference arm and the B200 arm
