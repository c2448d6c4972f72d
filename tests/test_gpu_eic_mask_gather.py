"""K2 (EIC update, thresholds, masks) and K3 (gather, bias compensation) vs the CPU oracle.

EIC scores, thresholds, masks, kept-channel index sets and gathered tensors are compared BIT-EXACT.
The bias-compensation GEMV is floating point with a different summation order than MKL's: tolerance
1e-5 * sum_i |act_i * W_sum[o,i]|.
"""
import numpy as np
import pytest
import torch

from oracle import eic_ref, gather_ref, mask_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _layers(seed, sizes):
    rng = np.random.RandomState(seed)
    return [rng.standard_normal(c).astype(np.float32) for c in sizes]


SIZES = [64, 64, 128, 256, 48, 512, 1024, 2048, 256, 33]


@pytest.mark.parametrize("r", [0.999, 0.99])
def test_eic_update_bit_exact(native, r):
    from dcfp_b200 import ops
    steps = 4
    gammas = [np.abs(g) + 0.5 for g in _layers(1, SIZES)]
    gammas[2][::3] *= -1  # negative gammas exercise the sign gate
    offs = np.concatenate([[0], np.cumsum(SIZES)]).astype(np.int32)
    eic = torch.empty(int(offs[-1]), dtype=torch.float32, device=DEV).fill_(float("nan"))  # first_step must ignore it
    gam_d = [torch.from_numpy(g).to(DEV) for g in gammas]
    offs_d = torch.from_numpy(offs).to(DEV)
    ref = [0] * len(SIZES)
    for t in range(steps):
        grads = _layers(10 + t, SIZES)
        for g in grads:
            g *= 1e-4
            g[::7] = 0.0  # exact zeros: gate false
        if t == 2:
            grads[0][1] = np.nan
            grads[0][2] = np.inf
            grads[1][:] = 1e-30  # grad*gamma underflows towards 0
            gammas_small = True
        ops.eic_update([torch.from_numpy(g).to(DEV) for g in grads], gam_d, offs_d, eic, r, first_step=(t == 0))
        ref = [eic_ref.eic_step(p, g, w, r) for p, g, w in zip(ref, grads, gammas)]
        got = eic.cpu().numpy()
        exp = np.concatenate(ref)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)) or np.array_equal(
            np.nan_to_num(got, nan=-1.0), np.nan_to_num(exp, nan=-1.0)), "step %d" % t
    # flat variant gives the same bits
    flat = torch.zeros(int(offs[-1]), dtype=torch.float32, device=DEV)
    g = torch.from_numpy(np.concatenate(_layers(99, SIZES))).to(DEV)
    w = torch.cat(gam_d)
    ops.eic_update_flat(g, w, flat, r, first_step=True)
    exp = eic_ref.eic_step(0, g.cpu().numpy(), w.cpu().numpy(), r)
    assert np.array_equal(flat.cpu().numpy().view(np.uint32), exp.view(np.uint32))


def _run_mask(ops, scores, groups, gp, layer_keep):
    sizes = [len(s) for s in scores]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=DEV)
    grp = torch.tensor(groups, dtype=torch.int32, device=DEV)
    mk = torch.tensor([mask_ref.min_keep_of(c, layer_keep) for c in sizes], dtype=torch.int32, device=DEV)
    k = []
    for g in (0, 1):
        n = sum(c for c, gg in zip(sizes, groups) if gg == g)
        k.append(mask_ref.thresh_index(n, gp) if n > 0 else -1)
    mask, thresh, kept = ops.thresh_mask(torch.from_numpy(np.concatenate(scores)).to(DEV), offs, grp, mk, k[0], k[1])
    return mask.cpu().numpy(), thresh.cpu().numpy(), kept.cpu().numpy()


@pytest.mark.parametrize("case", ["uniform", "eic_like", "ties", "one_group"])
def test_thresh_and_masks_bit_exact(native, case):
    from dcfp_b200 import ops
    rng = np.random.RandomState(5)
    sizes = [64, 64, 128, 256, 256, 512, 1024, 2048, 48, 256, 256, 512]
    groups = [0] * 8 + [1] * 4
    if case == "uniform":
        scores = [rng.rand(c).astype(np.float32) for c in sizes]
    elif case == "eic_like":  # O(1e-7) magnitudes, heavy-tailed
        scores = [(np.abs(rng.standard_normal(c)) ** 3 * 1e-7).astype(np.float32) for c in sizes]
    elif case == "ties":  # ~50% exact zeros, as after one EIC step, plus a layer that is all zero
        scores = [(rng.rand(c) * (rng.rand(c) > 0.5)).astype(np.float32) for c in sizes]
        scores[3][:] = 0.0
        scores[9][:] = scores[9][0]
    else:
        groups = [0] * len(sizes)
        scores = [rng.rand(c).astype(np.float32) for c in sizes]
    gp = 0.5
    while gp < 1.0:  # the exact float sequence prune.py walks (0.5, 0.52, 0.54, ...)
        mask, thresh, kept = _run_mask(ops, scores, groups, gp, 0.02)
        t_ref = mask_ref.thresholds(scores, groups, gp)
        m_ref = np.concatenate(mask_ref.masks(scores, groups, t_ref, 0.02))
        for g in (0, 1):
            assert np.float32(thresh[g]) == np.float32(t_ref[g]), (case, gp, g)
        assert np.array_equal(mask, m_ref), (case, gp)
        assert np.array_equal(np.nonzero(mask)[0], np.nonzero(m_ref)[0])
        offs = np.concatenate([[0], np.cumsum(sizes)])
        assert np.array_equal(kept, [int(m_ref[a:b].sum()) for a, b in zip(offs[:-1], offs[1:])])
        gp = gp + 0.02


def test_min_keep_fallback_tie_rule(native):
    from dcfp_b200 import ops
    # layer 1 sits entirely below the group threshold -> fallback keeps its top-5, ties lowest index first
    hi = np.linspace(1, 2, 1000).astype(np.float32)
    lo = np.zeros(256, dtype=np.float32)
    lo[[10, 20, 30]] = [0.3, 0.3, 0.2]
    mask, thresh, kept = _run_mask(ops, [hi, lo], [0, 0], 0.5, 0.02)
    m = mask[1000:]
    assert kept[1] == 5 and set(np.nonzero(m)[0]) == {10, 20, 30, 0, 1}


@pytest.mark.parametrize("shape", [(256, 128, 3, 3), (512, 1280, 1, 1), (19, 256, 1, 1), (64, 3, 3, 3), (2048,)])
def test_channel_gather_bit_exact(native, shape):
    from dcfp_b200 import ops
    rng = np.random.RandomState(3)
    W = rng.standard_normal(shape).astype(np.float32)
    out_idx = np.nonzero(rng.rand(shape[0]) > 0.5)[0].astype(np.int32)
    in_idx = np.nonzero(rng.rand(shape[1]) > 0.4)[0].astype(np.int32) if len(shape) > 1 else None
    Wd = torch.from_numpy(W).to(DEV)
    od = torch.from_numpy(out_idx).to(DEV)
    idd = None if in_idx is None else torch.from_numpy(in_idx).to(DEV)
    for oi, ii, o_np, i_np in [(od, idd, out_idx, in_idx), (od, None, out_idx, None), (None, idd, None, in_idx)]:
        if oi is None and ii is None:
            continue
        got = ops.channel_gather(Wd, oi, ii).cpu().numpy()
        exp = gather_ref.gather(W, o_np, i_np)
        assert got.shape == exp.shape and np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # empty selection
    empty = ops.channel_gather(Wd, torch.empty(0, dtype=torch.int32, device=DEV), None)
    assert empty.shape[0] == 0


def test_channel_gather_grouped_bit_exact(native):
    from dcfp_b200 import ops
    rng = np.random.RandomState(8)
    shapes = [(64, 3, 3, 3), (64,), (64,), (128, 64, 1, 1), (256, 128, 3, 3), (256,), (19, 256, 1, 1), (19,)]
    srcs, ois, iis, exp = [], [], [], []
    for s in shapes:
        W = rng.standard_normal(s).astype(np.float32)
        o = np.nonzero(rng.rand(s[0]) > 0.5)[0].astype(np.int32) if s[0] != 19 else None
        i = np.nonzero(rng.rand(s[1]) > 0.5)[0].astype(np.int32) if len(s) > 1 and s[1] != 3 else None
        srcs.append(torch.from_numpy(W).to(DEV))
        ois.append(None if o is None else torch.from_numpy(o).to(DEV))
        iis.append(None if i is None else torch.from_numpy(i).to(DEV))
        exp.append(gather_ref.gather(W, o, i))
    outs = ops.channel_gather_grouped(srcs, ois, iis)
    for got, e in zip(outs, exp):
        assert tuple(got.shape) == e.shape and np.array_equal(got.cpu().numpy().view(np.uint32), e.view(np.uint32))


@pytest.mark.parametrize("shape", [(512, 1280, 1, 1), (256, 560, 3, 3), (512, 4096, 3, 3)])
def test_bias_comp(native, shape):
    from dcfp_b200 import ops
    rng = np.random.RandomState(4)
    W = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    act = np.maximum(rng.standard_normal(shape[1]), 0).astype(np.float32) * (rng.rand(shape[1]) > 0.5)
    got = ops.bias_comp(torch.from_numpy(W).to(DEV), torch.from_numpy(act).to(DEV)).cpu().numpy()
    exp = gather_ref.bias_offset(W, act)
    mass = np.abs(W.astype(np.float64)).reshape(shape[0], shape[1], -1).sum(2) @ np.abs(act.astype(np.float64))
    assert (np.abs(got - exp) <= 1e-5 * mass + 1e-30).all()


def test_thresholds_large_score_vector_takes_the_layer_walk(native):
    """> ~170 k scores: the per-element group map no longer fits in shared memory -> the layer-by-layer select kernel."""
    from dcfp_b200 import ops
    rng = np.random.RandomState(17)
    sizes = [2048] * 100 + [333, 64]
    groups = [0] * 60 + [1] * 40 + [0, 1]
    scores = [(np.exp(rng.standard_normal(c)) * 1e-6).astype(np.float32) for c in sizes]
    for s in scores[::3]:
        s[rng.rand(s.size) < 0.3] = 0.0
    for gp in (0.5, 0.9):
        mask, thresh, kept = _run_mask(ops, scores, groups, gp, 0.02)
        t_ref = mask_ref.thresholds(scores, groups, gp)
        assert [float(t) for t in thresh] == [float(t) for t in t_ref]
        m_ref = np.concatenate(mask_ref.masks(scores, groups, t_ref, 0.02))
        assert np.array_equal(mask, m_ref)


@pytest.mark.parametrize("seed", list(range(24)))
def test_randomized_thresholds_masks_and_gathers_bit_exact(native, seed):
    """Seeded random layer tables (1..40 layers of 1..2048 channels, either group possibly empty), score distributions
    with heavy ties, global_percent from 0 to 0.999, layer_keep from 0 to 1 -- and
    random weight shapes / index lists (empty, single, all) through both gather entry points.  All bit-exact."""
    from dcfp_b200 import ops
    rng = np.random.RandomState(7000 + seed)
    n_layers = int(rng.randint(1, 41))
    sizes = [int(rng.choice([1, 2, 3, 7, 16, 33, 48, 64, 128, 255, 256, 512, 1024, 2048])) for _ in range(n_layers)]
    mode = seed % 4
    groups = [int(rng.randint(0, 2)) for _ in sizes] if mode != 3 else [1] * n_layers
    kind = rng.randint(0, 4)
    scores = []
    for c in sizes:
        if kind == 0:
            s = rng.rand(c)
        elif kind == 1:
            s = np.abs(rng.standard_normal(c)) ** 3 * 1e-7
        elif kind == 2:  # few distinct values: ties everywhere, including at the threshold
            s = rng.randint(0, 4, c) * 0.25
        else:
            s = rng.rand(c) * (rng.rand(c) > 0.6)
        scores.append(s.astype(np.float32))
    gp = float(rng.choice([0.0, 0.01, 0.3, 0.5, 0.52, 0.7, 0.9, 0.98, 0.999]))
    layer_keep = float(rng.choice([0.0, 0.01, 0.02, 0.1, 0.5, 1.0]))
    mask, thresh, kept = _run_mask(ops, scores, groups, gp, layer_keep)
    t_ref = mask_ref.thresholds(scores, groups, gp)
    m_ref = np.concatenate(mask_ref.masks(scores, groups, t_ref, layer_keep))
    for g in (0, 1):
        if any(gg == g for gg in groups):
            assert np.float32(thresh[g]) == np.float32(t_ref[g]), (g, gp)
    assert np.array_equal(mask, m_ref), (gp, layer_keep)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    assert np.array_equal(kept, [int(m_ref[a:b].sum()) for a, b in zip(offs[:-1], offs[1:])])

    # gathers: conv weights [O, I, kh, kw], BN vectors [O], depth-1 kernels, bf16-sized elements via float16 views
    srcs, ois, iis, exp = [], [], [], []
    for _ in range(int(rng.randint(1, 12))):
        O = int(rng.choice([1, 3, 19, 48, 64, 150, 256, 513]))
        if rng.rand() < 0.3:
            shape = (O,)
        else:
            khw = [(1, 1), (3, 3), (1, 7), (7, 7)][int(rng.randint(0, 4))]
            shape = (O, int(rng.choice([1, 3, 16, 64, 255, 320]))) + khw
        W = rng.standard_normal(shape).astype(np.float32)

        def pick(n):
            r = rng.rand()
            if r < 0.15:
                return None
            if r < 0.25:
                return np.arange(n, dtype=np.int32)
            if r < 0.35:
                return np.array([int(rng.randint(0, n))], dtype=np.int32)
            return np.nonzero(rng.rand(n) > rng.rand())[0].astype(np.int32)
        o = pick(shape[0])
        i = pick(shape[1]) if len(shape) > 1 else None
        srcs.append(torch.from_numpy(W).to(DEV))
        ois.append(None if o is None else torch.from_numpy(o).to(DEV))
        iis.append(None if i is None else torch.from_numpy(i).to(DEV))
        exp.append(gather_ref.gather(W, o, i))
    outs = ops.channel_gather_grouped(srcs, ois, iis)
    for got, e, s, o, i in zip(outs, exp, srcs, ois, iis):
        assert tuple(got.shape) == e.shape and np.array_equal(got.cpu().numpy().view(np.uint32), e.view(np.uint32))
        if o is not None or i is not None:
            one = ops.channel_gather(s, o, i)
            assert tuple(one.shape) == e.shape and np.array_equal(one.cpu().numpy().view(np.uint32), e.view(np.uint32))
