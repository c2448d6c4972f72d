"""Helpers shared by the CPU (oracle backend) and GPU (CUDA kernels) prune-parity tests: load a
tests/golden/prune_<cfg>.npz fixture written by the unmodified reference and compare a product run with it."""
import hashlib
import json
import os
import tempfile

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, "prune_%s.npz" % name))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def make_scores(model, kind, seed):
    """Same generator as tests/golden/make_golden.py (checked through the stored SHA-256)."""
    rng = np.random.RandomState(seed)
    eic = {}
    for n, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d) and n not in model.ignore_prune_layer:
            c = m.weight.numel()
            if kind == "uniform":
                s = rng.rand(c).astype(np.float32)
            else:
                s = (np.exp(rng.standard_normal(c) * 1.5) * 1e-7).astype(np.float32)
                s[rng.rand(c) < 0.4] = 0.0
            eic[n] = s
    return eic


def scores_digest(eic):
    h = hashlib.sha256()
    for n in eic:
        h.update(n.encode())
        h.update(np.ascontiguousarray(eic[n]).tobytes())
    return h.hexdigest()


def tensor_digest(t):
    t = t.detach().cpu().contiguous()
    return hashlib.sha256(str(tuple(t.shape)).encode() + str(t.dtype).encode() + t.numpy().tobytes()).hexdigest()


def build_model(cfg, beta_seed=None):
    from dcfp_b200.workloads.segnets import CONFIGS, build_segnet
    c = CONFIGS[cfg]
    model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0, with_loss=False)
    if beta_seed is not None:
        g = torch.Generator().manual_seed(beta_seed)
        with torch.no_grad():
            for mod in model.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.5)
    return model


def run_product_prune(model, eic, gp, layer_keep):
    """DCFPPruner.prune_model of the product (whatever backend torch.ops.dcfp / ops.* currently is)."""
    from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
    with tempfile.NamedTemporaryFile(suffix=".pth") as f:
        torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, f.name)
        pruner = DCFPPruner(global_percent=gp, layer_keep=layer_keep, score_file=f.name)
    sub, cfg = pruner.prune_model(model, except_start_keys=["conv_deepsup"])
    return pruner, sub, cfg


def check_case(z, meta, ci, pruner, sub, cfg, check_topology=True, skip=()):
    case = meta["cases"][ci]
    names = meta["module_names"]
    assert list(cfg.keys()) == names
    bits, counts = [], []
    for n in names:
        for side in ("in", "out"):
            if side + "_mask" in cfg[n]:
                m = np.asarray(cfg[n][side + "_mask"]).reshape(-1)
                assert cfg[n][side + "_mask"].dtype == np.float32 and cfg[n][side + "_mask"].ndim in (2, 4)
                bits.append(m.astype(np.uint8))
                counts.append((cfg[n][side + "_channels"], cfg[n]["raw_" + side + "_channels"]))
            else:
                counts.append((-1, -1))
    got_bits = np.packbits(np.concatenate(bits))
    assert np.array_equal(np.array(counts, dtype=np.int32), z["counts_%d" % ci]), "channel counts differ"
    assert np.array_equal(got_bits, z["masks_%d" % ci]), "kept-channel index sets differ"
    thresh = pruner._thresh if getattr(pruner, "_thresh", None) is not None else pruner.get_thresh()
    got_t = [int(np.float32(float(t)).view(np.uint32)) for t in thresh]
    assert got_t == case["thresh_bits"], "thresholds differ"
    if check_topology and ci == 0 and "norm_conv_links" in meta:
        assert [list(x) for x in pruner.norm_conv_links.items()] == meta["norm_conv_links"]
        assert list(pruner.except_layers) == meta["except_layers"]
        assert {k: set(v) for k, v in pruner.same_out_channel_groups.items()} == {k: set(v) for k, v in meta["groups"].items()}
        assert list(pruner.same_out_channel_groups.keys()) == list(meta["groups"].keys())
        assert list(pruner.modules_have_child) == meta["modules_have_child"]
        assert list(pruner.modules_have_ancest) == meta["modules_have_ancest"]
    sd = sub.state_dict()
    exp = case["state_dict_sha256"]
    keys = [k for k in sd if k not in skip]
    assert keys == list(exp.keys())
    bad = [k for k in keys if tensor_digest(sd[k]) != exp[k]]
    assert not bad, "pruned tensors not bit-exact: %s" % bad[:4]
    return sd


def load_sweep(name):
    z = np.load(os.path.join(GOLDEN, "sweep_%s.npz" % name))
    return z, json.loads(bytes(z["meta"]).decode())


def check_percent_sweep(cfg):
    """thresholds + raw keep masks of the product (whatever backend ops.* currently is) for all 25 global_percent
    values prune.py can visit, against the unmodified reference's (tests/golden/sweep_<cfg>.npz): bit-exact."""
    from dcfp_b200.pruners.channel_pruner import _structural_clone
    from dcfp_b200.pruners.dcfp_pruner import DCFPPruner
    z, meta = load_sweep(cfg)
    model = build_model(cfg)
    eic = make_scores(model, meta["kind"], meta["seed"])
    assert scores_digest(eic) == meta["scores_sha256"]
    with tempfile.NamedTemporaryFile(suffix=".pth") as f:
        torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, f.name)
        pruner = DCFPPruner(global_percent=0.5, layer_keep=meta["layer_keep"], score_file=f.name)
    clone = _structural_clone(model)
    pruner.prepare_from_supernet(clone)
    pruner.except_start_keys = pruner.except_start_keys + clone.ignore_prune_layer + ["conv_deepsup"]
    pruner.get_except_layers(clone)
    gp, n = 0.5, 0
    while gp < 1.0:  # prune.py:91,122 -- accumulated in floating point
        assert repr(gp) == meta["percents"][n]
        pruner.global_percent = gp
        thresh, masks = pruner._select()
        got_t = [int(np.float32(float(t)).view(np.uint32)) for t in thresh]
        assert got_t == meta["thresh_bits"][n], "thresholds differ at global_percent=%r" % gp
        bits = np.packbits(np.concatenate([masks[bn].numpy().reshape(-1).astype(np.uint8) for bn in meta["links"]]))
        assert np.array_equal(bits, z["masks_%d" % n]), "keep masks differ at global_percent=%r" % gp
        gp += 0.02
        n += 1
    assert n == len(meta["percents"]) == 25
