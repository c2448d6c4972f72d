"""Generates the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE (wzx99/DCFP).

Run in the build container only (needs /root/reference; oracle/ref_compat.py holds the three
arithmetic-free shims):

    python tests/golden/make_golden.py [eic] [prune] [scoring] [sweep] [balance]

The reference ships no tests or golden vectors, so these files ARE the pin of the oracle and of the
CUDA path on the GPU box, where /root/reference does not exist.

  eic_steps.npz        dcfp_pruning.step (pruners/dcfp_pruner.py:15-20) on stored grads / gammas
  prune_<cfg>.npz      DCFPPruner.prune_model (pruners/dcfp_pruner.py:43-92, channel_pruner.py:967-990)
                       on the BASELINE models c1..c4, random-init (seed 0), scores regenerated from a
                       numpy seed: thresholds, every in/out mask (bit-packed), channel counts, topology,
                       SHA-256 of every tensor of the pruned state_dict
  prune_c1_beta.npz    same with non-zero BN beta -> exercises bias compensation (channel_pruner.py:873-905);
                       stores the compensated running_mean vectors (fp32 GEMV: compared with a tolerance)
  sweep_c{1..4}.npz    thresholds + raw keep masks for all 25 global_percent values prune.py can visit
  balance.npz          BaseDataSet.get_label (datasets/Base.py:73-89): class-balance pixel weights, modes 1 and 2
  scoring_small.npz    reference Seg_Model + CriterionDSN + dcfp_pruning over 2 steps on 2x3x64x128 inputs:
                       per-step BN-gamma gradients and the final EIC (pins oracle/scoring_ref.py)
"""
import copy
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_compat  # noqa: E402

LAYER_KEEP = 0.02  # scripts/cs/prune.sh / prune.py:94


def percents():
    """global_percent values exactly as prune.py accumulates them (prune.py:91,122)."""
    out, gp = [], 0.5
    while gp < 1.0:
        out.append(gp)
        gp += 0.02
    return out


def make_scores(model, kind, seed):
    """Deterministic score vectors (numpy legacy MT19937 streams are stable across versions)."""
    rng = np.random.RandomState(seed)
    eic = {}
    for n, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d) and n not in model.ignore_prune_layer:
            c = m.weight.numel()
            if kind == "uniform":
                s = rng.rand(c).astype(np.float32)
            elif kind == "eic_like":  # O(1e-7) magnitudes, ~40 % exact zeros (sign gate failed every step)
                s = (np.exp(rng.standard_normal(c) * 1.5) * 1e-7).astype(np.float32)
                s[rng.rand(c) < 0.4] = 0.0
            else:
                raise ValueError(kind)
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


def build_ref_model(ref, cfg, beta_seed=None):
    from dcfp_b200.workloads.segnets import BACKBONE_PARA, CONFIGS
    c = CONFIGS[cfg]
    torch.manual_seed(0)
    model = getattr(ref.networks, c["arch"]).Seg_Model(backbone=c["backbone"], backbone_para=dict(BACKBONE_PARA), model_para={},
                                                       num_classes=c["num_classes"], align_corner=True, criterion=None, deepsup=True)
    if beta_seed is not None:
        g = torch.Generator().manual_seed(beta_seed)
        with torch.no_grad():
            for mod in model.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.5)
    return model


def gen_eic():
    ref = ref_compat.load_reference()
    sizes = [64, 64, 128, 256, 48, 512, 1024, 2048, 256, 33]
    rng = np.random.RandomState(7)
    steps = 5
    gammas = [(np.abs(rng.standard_normal(c)) + 0.5).astype(np.float32) for c in sizes]
    gammas[2][::3] *= -1
    gammas[4][5] = 0.0
    grads = []
    for t in range(steps):
        gs = [(rng.standard_normal(c) * 1e-4).astype(np.float32) for c in sizes]
        for g in gs:
            g[::7] = 0.0
        if t == 2:
            gs[0][1] = np.nan
            gs[0][2] = np.inf
            gs[1][:] = 1e-30
        grads.append(gs)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ignore_prune_layer = ["bn3"]
            for i, c in enumerate(sizes):
                setattr(self, "bn%d" % i, torch.nn.BatchNorm2d(c))

    out = dict(sizes=np.array(sizes), steps=steps)
    for i in range(len(sizes)):
        out["gamma_%d" % i] = gammas[i]
        for t in range(steps):
            out["grad_%d_%d" % (t, i)] = grads[t][i]
    for r in (0.999, 0.99):
        net = Net()
        tp = ref.dp.dcfp_pruning(net, r)
        assert "bn3" not in tp.state_dict["eic"]
        with torch.no_grad():
            for i in range(len(sizes)):
                getattr(net, "bn%d" % i).weight.copy_(torch.from_numpy(gammas[i]))
        for t in range(steps):
            for i in range(len(sizes)):
                getattr(net, "bn%d" % i).weight.grad = torch.from_numpy(grads[t][i].copy())
            tp.step(net)
            for i in range(len(sizes)):
                if i == 3:
                    continue
                v = tp.get_eic()["eic"]["bn%d" % i]
                assert v.dtype == torch.float32
                out["eic_r%s_%d_%d" % (str(r).replace(".", "p"), t, i)] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "eic_steps.npz"), **out)
    print("wrote eic_steps.npz")


def run_ref_prune(ref, model, eic, gp, tmp):
    torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, tmp)
    pruner = ref.dp.DCFPPruner(global_percent=gp, layer_keep=LAYER_KEEP, score_file=tmp)
    sub, cfg = pruner.prune_model(copy.deepcopy(model), except_start_keys=["conv_deepsup"])
    thresh = pruner.get_thresh()
    return pruner, sub, cfg, [float(t) for t in thresh]


def pack_cfg(cfg):
    names = list(cfg.keys())
    bits, counts = [], []
    for n in names:
        for side in ("in", "out"):
            if side + "_mask" in cfg[n]:
                m = np.asarray(cfg[n][side + "_mask"]).reshape(-1)
                assert set(np.unique(m)).issubset({0.0, 1.0})
                bits.append(m.astype(np.uint8))
                counts.append((cfg[n][side + "_channels"], cfg[n]["raw_" + side + "_channels"]))
            else:
                counts.append((-1, -1))
    return names, np.packbits(np.concatenate(bits)), np.array(counts, dtype=np.int32)


def gen_prune(only_beta=False):
    ref = ref_compat.load_reference()
    gps = percents()
    plan = {
        "c1": [("uniform", 1, gps[0]), ("uniform", 1, gps[10]), ("eic_like", 2, gps[0]), ("eic_like", 2, gps[20])],
        "c2": [("eic_like", 3, gps[0]), ("uniform", 4, gps[6])],
        "c3": [("eic_like", 5, gps[0]), ("uniform", 6, gps[6])],
        "c4": [("eic_like", 7, gps[0]), ("uniform", 8, gps[6])],
    }
    tmp = "/tmp/_golden_score.pth"
    for cfg_name, cases in ({} if only_beta else plan).items():
        model = build_ref_model(ref, cfg_name)
        out = {}
        meta = dict(cases=[], layer_keep=LAYER_KEEP)
        for ci, (kind, seed, gp) in enumerate(cases):
            eic = make_scores(model, kind, seed)
            pruner, sub, cfg, thresh = run_ref_prune(ref, model, eic, gp, tmp)
            names, packed, counts = pack_cfg(cfg)
            sd = sub.state_dict()
            meta["cases"].append(dict(kind=kind, seed=seed, global_percent=repr(gp), scores_sha256=scores_digest(eic),
                                      thresh_bits=[int(np.float32(t).view(np.uint32)) for t in thresh],
                                      state_dict_sha256={k: tensor_digest(v) for k, v in sd.items()}))
            out["masks_%d" % ci] = packed
            out["counts_%d" % ci] = counts
            if ci == 0:
                meta["module_names"] = names
                meta["norm_conv_links"] = list(pruner.norm_conv_links.items())
                meta["except_layers"] = list(pruner.except_layers)
                meta["groups"] = {k: list(v) for k, v in pruner.same_out_channel_groups.items()}
                meta["modules_have_child"] = list(pruner.modules_have_child)
                meta["modules_have_ancest"] = list(pruner.modules_have_ancest)
                meta["channel_spaces"] = [str(k) for k in pruner.channel_spaces.keys()]
            print(cfg_name, kind, gp, "thresh", thresh, "kept", int(counts[counts[:, 0] >= 0][1::2, 0].sum()))
        out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, "prune_%s.npz" % cfg_name), **out)
        print("wrote prune_%s.npz" % cfg_name)

    # non-zero beta: bias compensation moves running_mean of the consumers of pruned channels
    model = build_ref_model(ref, "c1", beta_seed=3)
    eic = make_scores(model, "uniform", 1)
    pruner, sub, cfg, thresh = run_ref_prune(ref, model, eic, gps[0], tmp)
    names, packed, counts = pack_cfg(cfg)
    sd = sub.state_dict()
    base = model.state_dict()
    out = dict(masks_0=packed, counts_0=counts)
    moved = []
    conv_bias = {n + ".bias" for n, m in sub.named_modules() if isinstance(m, torch.nn.Conv2d) and m.bias is not None}
    for k, v in sd.items():
        if k.endswith("running_mean") or k in conv_bias:  # the tensors resize_subnet_bias touches (fp32 GEMV)
            out["rm::" + k] = v.numpy().copy()
            moved.append(k)
    meta = dict(cases=[dict(kind="uniform", seed=1, global_percent=repr(gps[0]), beta_seed=3, scores_sha256=scores_digest(eic),
                            thresh_bits=[int(np.float32(t).view(np.uint32)) for t in thresh],
                            state_dict_sha256={k: tensor_digest(v) for k, v in sd.items() if k not in moved})],
                module_names=names, running_means=moved, layer_keep=LAYER_KEEP)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "prune_c1_beta.npz"), **out)
    print("wrote prune_c1_beta.npz; running_mean tensors:", len(moved), "of base", sum(k.endswith("running_mean") for k in base))


def gen_sweep(only=None):
    """get_thresh + gen_channel_mask (pruners/dcfp_pruner.py:43-92) for EVERY global_percent prune.py can visit
    (0.5, 0.52, ... accumulated in floating point exactly as prune.py:91,122 does), eic-like scores with 40 % exact
    zeros: thresholds (bits) and the raw per-link masks before propagation."""
    ref = ref_compat.load_reference()
    tmp = "/tmp/_golden_score.pth"
    for cfg_name, seed in (("c1", 11), ("c3", 12), ("c2", 13), ("c4", 14)):
        if only and cfg_name not in only:
            continue
        model = build_ref_model(ref, cfg_name)
        eic = make_scores(model, "eic_like", seed)
        torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, tmp)
        out, meta = {}, dict(kind="eic_like", seed=seed, scores_sha256=scores_digest(eic), layer_keep=LAYER_KEEP, percents=[], thresh_bits=[])
        pruner = None
        for gi, gp in enumerate(percents()):
            p = ref.dp.DCFPPruner(global_percent=gp, layer_keep=LAYER_KEEP, score_file=tmp)
            if pruner is None:  # trace once, reuse the topology for the other percents
                model_copy = copy.deepcopy(model)
                p.end_nodes = []  # channel_pruner.py:969-972 (set by prune_model)
                p.prepare_from_supernet(model_copy)
                p.except_start_keys = p.except_start_keys + model_copy.ignore_prune_layer + ["conv_deepsup"]
                p.get_except_layers(model_copy)
                pruner = p
            else:
                for attr in ("name2module", "module2name", "norm_conv_links", "conv_norm_links", "except_layers"):
                    setattr(p, attr, getattr(pruner, attr))
            thresh = p.get_thresh()
            p.gen_channel_mask()
            bits = []
            links = []
            for bn, conv in p.norm_conv_links.items():
                if conv not in p.except_layers:
                    bits.append(p.name2module[conv].out_mask.reshape(-1).numpy().astype(np.uint8))
                    links.append(bn)
            out["masks_%d" % gi] = np.packbits(np.concatenate(bits))
            meta["percents"].append(repr(gp))
            meta["thresh_bits"].append([int(np.float32(float(t)).view(np.uint32)) for t in thresh])
            meta["links"] = links
        out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, "sweep_%s.npz" % cfg_name), **out)
        print("wrote sweep_%s.npz (%d percents, %d links)" % (cfg_name, len(meta["percents"]), len(meta["links"])))


def gen_balance():
    """BaseDataSet.get_label (datasets/Base.py:73-89), the reference's own method called unbound on a stand-in `self`."""
    ref_compat.load_reference()
    import types
    import importlib.util  # `datasets` on sys.path is the Hugging Face package: load the reference's file by path
    spec = importlib.util.spec_from_file_location("dcfp_ref_datasets_base", os.path.join(ref_compat.REF_ROOT, "datasets", "Base.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    BaseDataSet = mod.BaseDataSet
    from dcfp_b200.workloads.synthetic import synthetic_labels
    out = {}
    cases = []
    for ci, (K, H, W, balance, beta) in enumerate([(19, 96, 160, 2, 0.9999), (19, 96, 160, 1, 0.9999), (150, 64, 64, 2, 0.9999),
                                                   (171, 80, 48, 2, 0.999)]):
        for img in range(3):
            label = synthetic_labels(100 * ci + img, K, H, W).numpy()
            present = np.unique(label[label != 255])
            cls = int(present[(7 * img) % len(present)])
            this = types.SimpleNamespace(balance=balance, ignore_label=255, num_classes=K, beta=beta)
            res = BaseDataSet.get_label(this, label, {"class": cls})
            assert res["weight"].dtype == np.float64 and np.array_equal(res["ori"], label)
            out["label_%d_%d" % (ci, img)] = label
            out["weight_%d_%d" % (ci, img)] = res["weight"]
            cases.append(dict(case=ci, img=img, K=K, balance=balance, beta=beta, sample_class=cls))
    out["meta"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "balance.npz"), **out)
    print("wrote balance.npz (%d images)" % len(cases))


def gen_scoring():
    """Reference model + loss + dcfp_pruning, the restated loop of train.py:255-268, tiny inputs."""
    ref = ref_compat.load_reference()
    from dcfp_b200.workloads.segnets import BACKBONE_PARA
    from dcfp_b200.workloads.synthetic import synthetic_batch
    K, H, W = 19, 64, 128
    torch.manual_seed(0)
    import types
    crit = ref.crit.CriterionDSN(dataset=types.SimpleNamespace(ignore_label=255))  # loss/criterion.py:52-60
    model = ref.networks.deeplabv3.Seg_Model(backbone="resnet50", backbone_para=dict(BACKBONE_PARA), model_para={}, num_classes=K,
                                             align_corner=True, criterion=crit, deepsup=True)
    model.train()
    tp = ref.dp.dcfp_pruning(model, 0.999)
    out = dict(K=K, H=H, W=W, steps=2)
    torch.set_num_threads(1)  # one thread: the summation order of the host convolutions is fixed
    for step in range(2):
        x, y = synthetic_batch([2 * step, 2 * step + 1], K, H, W)
        model.zero_grad()
        loss = model(x, y.long(), deepsup=True)
        loss["loss"].backward()
        tp.step(model)
        out["loss_%d" % step] = np.float32(loss["loss"].item())
        for n in tp.get_eic()["eic"]:
            out["grad_%d::%s" % (step, n)] = model.get_submodule(n).weight.grad.numpy().copy()
    for n, v in tp.get_eic()["eic"].items():
        out["eic::" + n] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "scoring_small.npz"), **out)
    print("wrote scoring_small.npz", len(tp.get_eic()["eic"]), "layers")


if __name__ == "__main__":
    what = sys.argv[1:] or ["eic", "prune", "scoring", "sweep", "balance"]
    assert ref_compat.available(), "reference tree not found"
    if "eic" in what:
        gen_eic()
    if "scoring" in what:
        gen_scoring()
    if "balance" in what:
        gen_balance()
    if "sweep" in what:
        gen_sweep([w for w in what if w in ("c1", "c2", "c3", "c4")])  # e.g. `make_golden.py sweep c2 c4`
    if "prune" in what or "beta" in what:
        gen_prune(only_beta="prune" not in what)
