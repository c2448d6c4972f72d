"""On-disk formats either side of the path: score.pth, channel_cfg.pth, pruned.pth.
Host logic only; the kernels are replaced by the CPU oracle through tests/fake_backend.py (test-only)."""
import os
import pickle

import numpy as np
import pytest
import torch

import golden_util as gu
from dcfp_b200.pruners import formats
from fake_backend import oracle_backend


@pytest.fixture(scope="module")
def pruned_c1():
    z, meta = gu.load_fixture("c1")
    case = meta["cases"][0]
    model = gu.build_model("c1")
    eic = gu.make_scores(model, case["kind"], case["seed"])
    with oracle_backend():
        _, sub, ccfg = gu.run_product_prune(model, eic, float(case["global_percent"]), meta["layer_keep"])
    return sub, ccfg, eic


def test_score_roundtrip_and_reference_layout(tmp_path):
    eic = {"backbone.bn1": torch.rand(64), "head.bn": torch.rand(7), "never_stepped": 0}
    p = str(tmp_path / "score.pth")
    formats.save_score(eic, p)
    raw = torch.load(p, map_location="cpu")  # what DCFPPruner.__init__ does (dcfp_pruner.py:34)
    assert list(raw.keys()) == ["eic"] and list(raw["eic"].keys()) == list(eic.keys())
    back = formats.load_score(p)
    assert back["never_stepped"] == 0 and isinstance(back["never_stepped"], int)
    assert all(torch.equal(back[k], eic[k]) for k in ("backbone.bn1", "head.bn"))
    eic["backbone.bn1"][0] = -1.0  # the file holds a copy, not a view
    assert formats.load_score(p)["backbone.bn1"][0] != -1.0
    torch.save({"not_eic": 1}, p)
    with pytest.raises(KeyError):
        formats.load_score(p)


def test_score_file_feeds_the_pruner(tmp_path):
    from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
    model = gu.build_model("c1")
    eic = {k: torch.from_numpy(v) for k, v in gu.make_scores(model, "uniform", 1).items()}
    p = str(tmp_path / "score.pth")
    formats.save_score(eic, p)
    pruner = DCFPPruner(global_percent=0.5, layer_keep=0.02, score_file=p)
    assert list(pruner.eic.keys()) == list(eic.keys())
    assert all(torch.equal(pruner.eic[k], eic[k]) for k in eic)


@pytest.mark.parametrize("portable", [False, True])
def test_channel_cfg_roundtrip(tmp_path, pruned_c1, portable):
    _, ccfg, _ = pruned_c1
    p = str(tmp_path / "channel_cfg.pth")
    formats.save_channel_cfg(ccfg, p, portable=portable)
    if portable:
        torch.load(p)  # default restricted unpickler of torch >= 2.6 accepts it
    else:
        with pytest.raises(pickle.UnpicklingError):
            torch.load(p)  # the failure prune.py:108 hits on a reference-written file
        ref_style = torch.load(p, weights_only=False)
        assert isinstance(ref_style[next(iter(ref_style))]["out_mask"], np.ndarray)
    back = formats.load_channel_cfg(p)
    assert list(back.keys()) == list(ccfg.keys())
    for name, cfg in ccfg.items():
        assert set(back[name].keys()) == set(cfg.keys())
        for k, v in cfg.items():
            if k.endswith("_mask"):
                assert isinstance(back[name][k], np.ndarray) and back[name][k].dtype == np.float32
                assert back[name][k].shape == v.shape and np.array_equal(back[name][k], v)
            else:
                assert type(back[name][k]) is int and back[name][k] == v


def test_channel_cfg_loader_stays_restricted(tmp_path):
    """Only the ndarray reconstruction globals are allow-listed: an arbitrary pickled object is still refused."""
    p = str(tmp_path / "evil.pth")
    torch.save({"conv": {"out_channels": 3, "out_mask": os.path.join}}, p)
    with pytest.raises(pickle.UnpicklingError):
        formats.load_channel_cfg(p)


@pytest.mark.parametrize("portable", [False, True])
def test_pruned_files_rebuild_the_subnet(tmp_path, pruned_c1, portable):
    """prune.py:97-110: two files -> fresh model -> init_pruned_model -> load -> same outputs, bit for bit."""
    sub, ccfg, _ = pruned_c1
    weights, cfg = formats.save_pruned(sub, ccfg, str(tmp_path / "out"), portable=portable)
    assert os.path.basename(weights) == "pruned.pth" and os.path.basename(cfg) == "channel_cfg.pth"
    fresh = gu.build_model("c1")
    formats.load_pruned_model(fresh, cfg, weights)
    sd, sd2 = sub.state_dict(), fresh.state_dict()
    assert list(sd.keys()) == list(sd2.keys()) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    for (n, m), (_, m2) in zip(sub.named_modules(), fresh.named_modules()):
        if isinstance(m, torch.nn.Conv2d):
            assert (m.in_channels, m.out_channels, m.groups) == (m2.in_channels, m2.out_channels, m2.groups), n
    sub.eval()
    fresh.eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        a, b = sub(x, deepsup=True), fresh(x, deepsup=True)
    assert all(torch.equal(i, j) for i, j in zip(a, b))


def test_load_state_envelopes_prefixes_and_report(tmp_path):
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    sd = {k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in net.state_dict().items()}
    for envelope in (None, "model", "state_dict"):
        p = str(tmp_path / "w.pth")
        torch.save(sd if envelope is None else {envelope: sd, "iteration": 7}, p)
        tgt = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
        assert formats.load_state(tgt, p) == ([], [])
        assert all(torch.equal(tgt.state_dict()[k], sd[k]) for k in sd)
    # DataParallel-style prefix stripped (pyt_utils.py:56-61)
    tgt = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    assert formats.load_state(tgt, {"module." + k: v for k, v in sd.items()}, ignore_prefix="module.") == ([], [])
    assert torch.equal(tgt[0].weight, sd["0.weight"])
    # prefix added (:63-68)
    wrapped = torch.nn.ModuleDict({"net": torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))})
    assert formats.load_state(wrapped, sd, extra_prefix="net.") == ([], [])
    assert torch.equal(wrapped["net"][0].weight, sd["0.weight"])
    # the report leaves num_batches_tracked out, lists the rest (:70-90)
    partial = {k: v for k, v in sd.items() if k not in ("1.num_batches_tracked", "0.bias")}
    partial["2.weight"] = torch.zeros(1)
    tgt = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    assert formats.load_state(tgt, partial) == (["0.bias"], ["2.weight"])
    with pytest.raises(RuntimeError, match="0.bias"):
        formats.load_pruned_model(torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4)), {}, partial)


@pytest.mark.ref
def test_files_cross_read_with_unmodified_reference(tmp_path, pruned_c1):
    """Files written here are consumed by the reference's own init_pruned_model + load_model (prune.py:108-110),
    and a channel_cfg the reference's export_subnet wrote is read back identically by load_channel_cfg."""
    from oracle import ref_compat
    ref = ref_compat.load_reference()
    sub, ccfg, eic = pruned_c1
    weights, cfg = formats.save_pruned(sub, ccfg, str(tmp_path / "out"))
    fresh = gu.build_model("c1")
    ref.cp.init_pruned_model(fresh, torch.load(cfg, weights_only=False))
    import importlib
    try:  # REF_ROOT is on sys.path after load_reference(); `utils` is the reference's package
        pyt = importlib.import_module("utils.pyt_utils")
        assert pyt.__file__.startswith(ref_compat.REF_ROOT), pyt.__file__
    except ImportError as e:  # its logger / distributed imports are outside the path
        pytest.skip("reference utils/pyt_utils.py does not import here: %r" % (e,))
    pyt.load_model(fresh, weights)
    sd, sd2 = sub.state_dict(), fresh.state_dict()
    assert list(sd.keys()) == list(sd2.keys()) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    # the reverse direction: the reference prunes and writes, this side reads
    import copy
    model = gu.build_model("c1")
    score = str(tmp_path / "score.pth")
    formats.save_score({k: torch.from_numpy(v) for k, v in eic.items()}, score)
    rp = ref.dp.DCFPPruner(global_percent=0.5, layer_keep=0.02, score_file=score)
    _, ref_cfg = rp.prune_model(copy.deepcopy(model), except_start_keys=["conv_deepsup"])
    p = str(tmp_path / "ref_cfg.pth")
    torch.save(ref_cfg, p)
    back = formats.load_channel_cfg(p)
    assert list(back.keys()) == list(ref_cfg.keys())
    for n in ref_cfg:
        for k, v in ref_cfg[n].items():
            assert np.array_equal(np.asarray(back[n][k]), np.asarray(v)), (n, k)
