"""K1 parity: CUDA `class_stats` (through torch.ops.dcfp -> C ABI) vs the CPU oracle.

Tolerance (fp32 accumulation inside a CTA, fp64 across CTAs):
    |S - S_ref| <= RTOL * sum|v|  per (class, channel)     with RTOL = 1e-5
counts are integers and must match exactly.
"""
import pytest
import torch

from oracle import class_stats_ref as ref

RTOL = 1e-5
pytestmark = pytest.mark.gpu


def _labels(n, h0, w0, K, dtype, seed, blob=True):
    g = torch.Generator().manual_seed(seed)
    if blob:  # coarse random blocks -> coherent regions, then ~3% ignore
        coarse = torch.randint(0, K, (n, max(h0 // 16, 1), max(w0 // 16, 1)), generator=g)
        lab = torch.nn.functional.interpolate(coarse[:, None].float(), size=(h0, w0), mode="nearest")[:, 0].long()
    else:
        lab = torch.randint(0, K, (n, h0, w0), generator=g)
    ign = torch.rand(n, h0, w0, generator=g) < 0.03
    lab[ign] = 255
    return lab.to(dtype)


def _keys(ops, label, h, w, K, cnt=None):
    """label keys through the CUDA path, checked bit-exact against the oracle's nearest labels."""
    keys = ops.label_keys(label.to("cuda"), h, w, K, cnt)
    lab = ref.nearest_labels(label, h, w)
    exp = torch.where((lab >= 0) & (lab < K), lab, torch.full_like(lab, K)).to(torch.uint8)
    assert torch.equal(keys.cpu(), exp)
    return keys


def _check(ops, x, label, K, dy=None, scale=None, shift=None, cnt=True):
    dev = torch.device("cuda")
    C, h, w = x.shape[1:]
    S1 = torch.zeros(K, C, dtype=torch.float64, device=dev)
    S2 = torch.zeros_like(S1)
    cn = torch.zeros(K, dtype=torch.float64, device=dev) if cnt else None
    to = lambda t: None if t is None else t.to(dev)
    xd = x.to(dev)
    dyd = to(dy)
    if dyd is not None and not x.is_contiguous():
        dyd = dyd.contiguous(memory_format=torch.channels_last)
    keys = _keys(ops, label, h, w, K, cn)
    ops.class_stats(xd, keys, K, S1, S2, dy=dyd, scale=to(scale), shift=to(shift))
    torch.cuda.synchronize()
    v = ref.functor_fwd(x.float(), scale, shift) if dy is None else ref.functor_bwd(x.float(), dy.float(), scale, shift)
    rc, r1, r2 = ref.class_stats(v, label, K)
    mass = ref.abs_mass(v, label, K)
    assert torch.equal(cn.cpu(), rc) if cnt else True
    e1 = (S1.cpu() - r1).abs()
    assert (e1 <= RTOL * mass + 1e-30).all(), "S1 max rel-to-mass err %.3g" % (e1 / (mass + 1e-30)).max()
    e2 = (S2.cpu() - r2).abs()
    assert (e2 <= RTOL * r2 + 1e-30).all(), "S2 max rel err %.3g" % (e2 / (r2 + 1e-30)).max()
    return S1, S2, cn


SHAPES = [
    # N, C, h, w, H0, W0, K
    (2, 256, 64, 128, 512, 1024, 19),   # the dominant c1/c2 shape
    (2, 64, 128, 256, 512, 1024, 19),
    (2, 48, 128, 128, 512, 512, 171),   # partial channel group, shared-atomic accumulators
    (1, 33, 20, 36, 160, 288, 19),      # ragged: plane = 11 full segments + 16 px
    (2, 512, 6, 6, 512, 512, 150),      # PSP pyramid stage (generic path)
    (2, 256, 1, 1, 512, 1024, 19),      # ASPP image pooling (generic path)
    (2, 64, 97, 97, 769, 769, 19),      # odd crop, non-integer label ratio (generic path)
    (3, 96, 40, 52, 300, 411, 150),     # non-integer ratio on the tiled path
    (2, 32, 64, 64, 64, 64, 2),         # label at feature resolution
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fwd_matches_oracle(native, shape, dtype):
    from dcfp_b200 import ops
    N, C, h, w, H0, W0, K = shape
    g = torch.Generator().manual_seed(hash(shape) % 2**31)
    x = (torch.randn(N, C, h, w, generator=g) * 1.5 + 0.3).to(dtype)
    label = _labels(N, H0, W0, K, torch.uint8, seed=7 + C)
    _check(ops, x, label, K)


@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int32, torch.int64])
def test_label_dtypes_and_affine(native, label_dtype):
    from dcfp_b200 import ops
    N, C, h, w, K = 2, 128, 64, 64, 19
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, C, h, w, generator=g)
    scale = torch.rand(C, generator=g) + 0.5
    shift = torch.randn(C, generator=g)
    label = _labels(N, 512, 512, K, label_dtype, seed=3)
    _check(ops, x, label, K, scale=scale, shift=shift)


def test_iid_labels_worst_case_runs(native):
    from dcfp_b200 import ops
    N, C, h, w, K = 2, 64, 64, 128, 19
    x = torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(5))
    label = _labels(N, h, w, K, torch.uint8, seed=9, blob=False)
    _check(ops, x, label, K)


def test_all_ignored_and_single_class(native):
    from dcfp_b200 import ops
    x = torch.randn(2, 64, 32, 64, generator=torch.Generator().manual_seed(1))
    lab = torch.full((2, 256, 512), 255, dtype=torch.uint8)
    S1, S2, cn = _check(ops, x, lab, 19)
    assert S1.abs().sum() == 0 and cn.sum() == 0
    _check(ops, x, torch.zeros(2, 256, 512, dtype=torch.uint8), 19)
    # K == 1 without labels: plain per-channel sums
    dev = torch.device("cuda")
    S1 = torch.zeros(1, 64, dtype=torch.float64, device=dev)
    S2 = torch.zeros_like(S1)
    ops.class_stats(x.to(dev), None, 1, S1, S2)
    torch.cuda.synchronize()
    assert ((S1[0].cpu() - x.double().sum((0, 2, 3))).abs() <= RTOL * x.double().abs().sum((0, 2, 3))).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bwd_functor(native, dtype):
    from dcfp_b200 import ops
    N, C, h, w, K = 2, 256, 64, 128, 19
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(N, C, h, w, generator=g) * 2 + 1).to(dtype)
    dy = (torch.randn(N, C, h, w, generator=g) * 1e-3).to(dtype)
    mean = x.float().mean((0, 2, 3))
    invstd = 1.0 / torch.sqrt(x.float().var((0, 2, 3), unbiased=False) + 1e-5)
    label = _labels(N, 512, 1024, K, torch.int64, seed=4)
    _check(ops, x, label, K, dy=dy, scale=invstd, shift=-mean * invstd)


def test_channels_last(native):
    from dcfp_b200 import ops
    N, C, h, w, K = 2, 96, 32, 64, 19
    x = torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(2)).contiguous(memory_format=torch.channels_last)
    label = _labels(N, 256, 512, K, torch.uint8, seed=8)
    _check(ops, x, label, K)


def test_grouped_equals_per_layer_and_accumulates(native):
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    K = 19
    g = torch.Generator().manual_seed(33)
    shapes = [(2, 64, 64, 128), (2, 256, 64, 128), (2, 128, 128, 256), (2, 256, 1, 1), (2, 48, 32, 64)]
    xs = [torch.randn(*s, generator=g) for s in shapes]
    label = _labels(2, 512, 1024, K, torch.uint8, seed=12)
    xd = [x.to(dev) for x in xs]
    keys = {}
    for s in shapes:  # one key plane per label resolution, shared by the layers of that resolution
        if s[2:] not in keys:
            keys[s[2:]] = _keys(ops, label, s[2], s[3], K)
    kl = [keys[s[2:]] for s in shapes]
    S1 = [torch.zeros(K, s[1], dtype=torch.float64, device=dev) for s in shapes]
    S2 = [torch.zeros_like(t) for t in S1]
    ops.class_stats_grouped(xd, kl, K, S1, S2)
    ops.class_stats_grouped(xd, kl, K, S1, S2)  # += semantics: second pass doubles
    torch.cuda.synchronize()
    for x, a1, a2 in zip(xs, S1, S2):
        rc, r1, r2 = ref.class_stats(x, label, K)
        mass = ref.abs_mass(x, label, K)
        assert ((a1.cpu() - 2 * r1).abs() <= 2 * RTOL * mass + 1e-30).all()
        assert ((a2.cpu() - 2 * r2).abs() <= 2 * RTOL * r2 + 1e-30).all()


def test_dgamma_equals_autograd(native):
    """sum_k S1_bwd[k, c] must equal autograd's bn.weight.grad (the quantity dcfp_pruner.py:18 reads)."""
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(0)
    N, C, h, w, K = 2, 256, 64, 128, 19
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    x = (torch.randn(N, C, h, w, device=dev) * 3 + 1).requires_grad_(True)
    y = bn(x)
    dy = torch.randn_like(y) * 1e-3
    y.backward(dy)
    mean = x.detach().mean((0, 2, 3))
    invstd = 1.0 / torch.sqrt(x.detach().var((0, 2, 3), unbiased=False) + bn.eps)
    label = _labels(N, 512, 1024, K, torch.int64, seed=6).to(dev)
    label[label == 255] = 0  # dgamma sums over every pixel: no ignored ones here
    S1 = torch.zeros(K, C, dtype=torch.float64, device=dev)
    S2 = torch.zeros_like(S1)
    keys = ops.label_keys(label, h, w, K)
    ops.class_stats(x.detach(), keys, K, S1, S2, dy=dy, scale=invstd, shift=-mean * invstd)
    dgamma = ops.reduce_classes(S1)
    ref_g = bn.weight.grad
    tol = 1e-5 * ref_g.abs() + 1e-5 * ref_g.abs().mean()
    assert ((dgamma - ref_g).abs() <= tol).all(), ((dgamma - ref_g).abs() / tol).max()


def test_validation_errors(native):
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    x = torch.randn(1, 8, 4, 4, device=dev)
    S = torch.zeros(300, 8, dtype=torch.float64, device=dev)
    with pytest.raises(RuntimeError, match="K=300"):
        ops.class_stats(x, torch.zeros(1, 4, 4, dtype=torch.uint8, device=dev), 300, S, S.clone())
    with pytest.raises(RuntimeError, match="K=300"):
        ops.label_keys(torch.zeros(1, 4, 4, dtype=torch.uint8, device=dev), 4, 4, 300)
    with pytest.raises(RuntimeError, match="keys must be"):
        ops.class_stats(x, torch.zeros(1, 8, 8, dtype=torch.uint8, device=dev), 19, S[:19], S[:19].clone())
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.class_stats(x.cpu(), None, 1, S[:1], S[:1].clone())


NHWC_SHAPES = [
    # N, C, h, w, H0, W0, K -- channels_last feature maps (the layout bench.py scores: cuDNN's NHWC-native convolutions)
    (2, 256, 64, 128, 512, 1024, 19),   # 2 slabs x 4 pixel phases
    (2, 64, 128, 256, 512, 1024, 19),   # one slab, half of the lanes idle
    (2, 48, 32, 64, 256, 512, 19),      # partial slab
    (2, 1024, 32, 64, 256, 512, 19),    # 8 slabs, one phase
    (1, 2048, 16, 32, 128, 256, 19),    # two slab groups
    (2, 384, 24, 40, 192, 320, 7),      # 3 slabs -> spc 4, one slab of the group is empty
    (3, 128, 31, 37, 250, 300, 19),     # ragged: n_px not a multiple of the group / chunk
    (2, 100, 20, 20, 160, 160, 19),     # C % 4 == 0 but not a multiple of the slab
    (2, 66, 20, 20, 160, 160, 19),      # C % 4 != 0 -> generic path
    (2, 128, 32, 32, 256, 256, 150),    # K = 150 > 12 slot rows: tagged slot cache with evictions (any K <= 255 stays on the TMA path)
    (2, 512, 2, 2, 512, 512, 19),       # tiny pooled map -> generic path
]


@pytest.mark.parametrize("shape", NHWC_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("bwd", [False, True])
def test_channels_last_fast_path(native, shape, dtype, bwd):
    from dcfp_b200 import ops
    N, C, h, w, H0, W0, K = shape
    g = torch.Generator().manual_seed((hash(shape) + bwd) % 2**31)
    x = (torch.randn(N, C, h, w, generator=g) * 1.5 + 0.3).to(dtype).contiguous(memory_format=torch.channels_last)
    label = _labels(N, H0, W0, K, torch.uint8, seed=17 + C)
    if bwd:
        dy = (torch.randn(N, C, h, w, generator=g) * 1e-3).to(dtype).contiguous(memory_format=torch.channels_last)
        mean = x.float().mean((0, 2, 3))
        invstd = 1.0 / torch.sqrt(x.float().var((0, 2, 3), unbiased=False) + 1e-5)
        _check(ops, x, label, K, dy=dy, scale=invstd, shift=-mean * invstd)
    else:
        _check(ops, x, label, K)


def test_channels_last_invstd_mean_mode_and_no_keys(native):
    """affine_mode = INVSTD_MEAN (what the scorer passes: autograd's saved mean / invstd) and keys == None (K == 1)."""
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    N, C, h, w, K = 2, 256, 32, 64, 19
    g = torch.Generator().manual_seed(41)
    for layout in (torch.contiguous_format, torch.channels_last):
        x = (torch.randn(N, C, h, w, generator=g) * 2 + 1).contiguous(memory_format=layout)
        dy = (torch.randn(N, C, h, w, generator=g) * 1e-3).contiguous(memory_format=layout)
        mean = x.mean((0, 2, 3))
        invstd = 1.0 / torch.sqrt(x.var((0, 2, 3), unbiased=False) + 1e-5)
        label = _labels(N, 256, 512, K, torch.uint8, seed=5)
        keys = ops.label_keys(label.to(dev), h, w, K)
        S1 = torch.zeros(K, C, dtype=torch.float64, device=dev)
        S2 = torch.zeros_like(S1)
        ops.class_stats(x.to(dev), keys, K, S1, S2, dy=dy.to(dev), scale=invstd.to(dev), shift=mean.to(dev),
                        affine_mode=ops.AFFINE_INVSTD_MEAN)
        rc, r1, r2 = ref.class_stats_bwd(x, dy, mean, invstd, label, K)
        mass = ref.abs_mass(ref.functor_bwd(x, dy, invstd, -mean * invstd), label, K)
        assert ((S1.cpu() - r1).abs() <= 2 * RTOL * mass + 1e-30).all()
        assert ((S2.cpu() - r2).abs() <= 2 * RTOL * r2 + 1e-30).all()
        T1 = torch.zeros(1, C, dtype=torch.float64, device=dev)
        T2 = torch.zeros_like(T1)
        ops.class_stats(x.to(dev), None, 1, T1, T2)
        assert ((T1[0].cpu() - x.double().sum((0, 2, 3))).abs() <= RTOL * x.double().abs().sum((0, 2, 3))).all()
        assert ((T2[0].cpu() - (x.double() ** 2).sum((0, 2, 3))).abs() <= RTOL * (x.double() ** 2).sum((0, 2, 3))).all()


@pytest.mark.parametrize("K", [5, 19, 171])
def test_channels_last_iid_labels_thrash_the_slot_cache(native, K):
    """i.i.d. labels: every pixel opens a new run and (K > 8) almost every run evicts a row of the per-warp slot cache."""
    from dcfp_b200 import ops
    N, C, h, w = 2, 192, 24, 40
    x = torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(K)).contiguous(memory_format=torch.channels_last)
    label = _labels(N, h, w, K, torch.uint8, seed=K + 1, blob=False)
    _check(ops, x, label, K)


@pytest.mark.parametrize("shape", [(2, 64, 24, 40, 19, False), (2, 64, 24, 40, 150, False), (3, 64, 5, 7, 19, True),
                                   (2, 64, 64, 128, 19, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("bwd", [False, True])
def test_channels_last_64_channel_pixel_pair_rows(native, shape, dtype, bwd):
    """C == 64: the NHWC kernel views the map as [n_px / 2][128] (a row = 2 pixels, lanes 16-31 hold the odd pixel).
    i.i.d. labels make almost every row straddle two classes; an odd pixel count falls back to single-pixel rows."""
    from dcfp_b200 import ops
    N, C, h, w, K, blob = shape
    g = torch.Generator().manual_seed(h * 131 + K + bwd)
    x = (torch.randn(N, C, h, w, generator=g) * 1.5 + 0.3).to(dtype).contiguous(memory_format=torch.channels_last)
    label = _labels(N, h * (8 if blob else 1), w * (8 if blob else 1), K, torch.uint8, seed=K + h, blob=blob)
    if bwd:
        dy = (torch.randn(N, C, h, w, generator=g) * 1e-3).to(dtype).contiguous(memory_format=torch.channels_last)
        mean = x.float().mean((0, 2, 3))
        invstd = 1.0 / torch.sqrt(x.float().var((0, 2, 3), unbiased=False) + 1e-5)
        _check(ops, x, label, K, dy=dy, scale=invstd, shift=-mean * invstd)
    else:
        _check(ops, x, label, K)


@pytest.mark.parametrize("layout", ["nchw", "nhwc", "nhwc64", "generic"])
@pytest.mark.parametrize("K", [19, 150])
def test_shared_arena_columns_and_guards(native, layout, K):
    """Layers write into column slices of ONE [K, sumC] arena (row stride ld): the neighbouring columns, pre-filled
    with sentinels, must come back untouched (compute-sanitizer is closed on this pool: own guards instead), and a
    second call accumulates on top of the first."""
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    shape = {"nchw": (2, 96, 32, 64), "nhwc": (2, 160, 24, 40), "nhwc64": (2, 64, 24, 40), "generic": (2, 30, 5, 7)}[layout]
    N, C, h, w = shape
    g = torch.Generator().manual_seed(C + K)
    x = torch.randn(N, C, h, w, generator=g)
    if layout.startswith("nhwc"):
        x = x.contiguous(memory_format=torch.channels_last)
    label = _labels(N, h * 4, w * 4, K, torch.uint8, seed=3 * K)
    keys = ops.label_keys(label.to(dev), h, w, K)
    left, right = 37, 11
    arena = torch.full((2, K, left + C + right), 12345.0, dtype=torch.float64, device=dev)
    S1 = arena[0][:, left:left + C]
    S2 = arena[1][:, left:left + C]
    S1.zero_()
    S2.zero_()
    ops.class_stats(x.to(dev), keys, K, S1, S2)
    ops.class_stats(x.to(dev), keys, K, S1, S2)
    torch.cuda.synchronize()
    assert (arena[:, :, :left] == 12345.0).all() and (arena[:, :, left + C:] == 12345.0).all(), "wrote outside its columns"
    rc, r1, r2 = ref.class_stats(x, label, K)
    mass = ref.abs_mass(x, label, K)
    assert ((S1.cpu() - 2 * r1).abs() <= 2 * RTOL * mass + 1e-30).all()
    assert ((S2.cpu() - 2 * r2).abs() <= 2 * RTOL * r2 + 1e-30).all()


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_maximum_class_count(native, layout):
    """K = 255 = DCFP_MAX_CLASSES (uint8 keys, 255 is the largest usable class count); label 255 itself is class 254's
    neighbour 'ignore' only when K < 256 -> with K = 255 every label in [0, 255) is a class and 255 is dropped."""
    from dcfp_b200 import ops
    N, C, h, w, K = 2, 128, 32, 32, 255
    x = torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(255))
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    label = torch.randint(0, 256, (N, h, w), generator=torch.Generator().manual_seed(7), dtype=torch.int64).to(torch.uint8)
    _check(ops, x, label, K)


def test_more_layers_than_one_launch_holds(native):
    """A grouped call with more layers than one launch's parameter space holds (160 / 128 / 96 / 80 per path) is split
    transparently; every layer still lands in its own arena columns."""
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    K, n_layers, C = 19, 171, 32
    g = torch.Generator().manual_seed(5)
    label = _labels(2, 128, 128, K, torch.uint8, seed=2)
    keys = ops.label_keys(label.to(dev), 16, 16, K)
    for fmt in (torch.contiguous_format, torch.channels_last):
        xs = [torch.randn(2, C, 16, 16, generator=g).contiguous(memory_format=fmt) for _ in range(n_layers)]
        arena = torch.zeros(2, K, n_layers * C, dtype=torch.float64, device=dev)
        S1 = [arena[0][:, i * C:(i + 1) * C] for i in range(n_layers)]
        S2 = [arena[1][:, i * C:(i + 1) * C] for i in range(n_layers)]
        xd = [x.to(dev) for x in xs]
        ops.class_stats_grouped(xd, [keys] * n_layers, K, S1, S2)
        dys = [torch.randn_like(x) for x in xd]
        sc = [torch.ones(C, device=dev)] * n_layers
        sf = [torch.zeros(C, device=dev)] * n_layers
        arena_b = torch.zeros_like(arena)
        ops.class_stats_grouped(xd, [keys] * n_layers, K, [arena_b[0][:, i * C:(i + 1) * C] for i in range(n_layers)],
                                [arena_b[1][:, i * C:(i + 1) * C] for i in range(n_layers)], dys=dys, scales=sc, shifts=sf)
        torch.cuda.synchronize()
        for i in (0, 79, 80, 95, 96, 127, 128, 159, 160, 170):
            rc, r1, r2 = ref.class_stats(xs[i], label, K)
            mass = ref.abs_mass(xs[i], label, K)
            assert ((S1[i].cpu() - r1).abs() <= RTOL * mass + 1e-30).all(), i
            v = ref.functor_bwd(xs[i], dys[i].cpu(), torch.ones(C), torch.zeros(C))
            _, b1, _ = ref.class_stats(v, label, K)
            assert ((arena_b[0][:, i * C:(i + 1) * C].cpu() - b1).abs() <= RTOL * ref.abs_mass(v, label, K) + 1e-30).all(), i


def test_empty_and_bad_inputs_are_rejected(native):
    from dcfp_b200 import ops
    dev = torch.device("cuda")
    S = torch.zeros(19, 8, dtype=torch.float64, device=dev)
    with pytest.raises(RuntimeError, match="non-null|bad extent"):
        ops.class_stats(torch.empty(0, 8, 4, 4, device=dev), torch.empty(0, 4, 4, dtype=torch.uint8, device=dev), 19, S, S.clone())
    with pytest.raises(RuntimeError, match="fp32 or bf16"):
        ops.class_stats(torch.zeros(1, 8, 4, 4, device=dev, dtype=torch.float16), torch.zeros(1, 4, 4, dtype=torch.uint8, device=dev),
                        19, S, S.clone())
    with pytest.raises(RuntimeError, match="empty layer list"):
        ops.class_stats_grouped([], [], 19, [], [])


@pytest.mark.parametrize("K", [40, 171, 255])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nchw_large_k_remap_overflow(native, K, dtype):
    """NCHW, K > 24: i.i.d. labels put far more than 32 classes into every tile -> the classes beyond the 32 remapped
    rows take the direct-to-arena path."""
    from dcfp_b200 import ops
    N, C, h, w = 2, 64, 32, 64
    x = (torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(K)) * 1.3).to(dtype)
    label = torch.randint(0, K, (N, h, w), generator=torch.Generator().manual_seed(K + 1)).to(torch.uint8)
    _check(ops, x, label, K)
    dy = (torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(K + 2)) * 1e-2).to(dtype)
    _check(ops, x, label, K, dy=dy, scale=torch.rand(C) + 0.5, shift=torch.randn(C))


def _fuzz_cases(n_cases=160, seed=20240607):
    """Seeded random shapes across every dispatch boundary of K1: channel counts around the 4 / 32 / 64 / 128 edges,
    pixel counts around the TMA box sizes (ragged tails, planes shorter than one box, odd pixel counts under the
    64-channel pixel-pair view), K around the slot-cache (12/13), table (24/25) and key (255) limits."""
    import random
    rnd = random.Random(seed)
    Cs = [4, 8, 12, 28, 32, 36, 48, 60, 64, 68, 96, 124, 128, 132, 192, 256, 260, 320, 33, 7]
    Ks = [1, 2, 11, 12, 13, 19, 23, 24, 25, 26, 60, 150, 171, 254, 255]
    cases = []
    for i in range(n_cases):
        C = rnd.choice(Cs)
        K = rnd.choice(Ks)
        N = rnd.choice([1, 2, 3])
        h = rnd.choice([1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 40, 64])
        w = rnd.choice([1, 2, 4, 7, 8, 9, 16, 24, 31, 32, 33, 48, 65, 128])
        layout = rnd.choice(["nchw", "nhwc"])
        dtype = rnd.choice([torch.float32, torch.float32, torch.bfloat16])
        functor = rnd.choice(["fwd", "fwd_affine", "bwd", "bwd"])
        blob = rnd.random() < 0.7
        cases.append((i, N, C, h, w, K, layout, dtype, functor, blob))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: "%d-%dx%dx%dx%d-K%d-%s-%s-%s" % (
    c[0], c[1], c[2], c[3], c[4], c[5], c[6], "bf16" if c[7] == torch.bfloat16 else "f32", c[8]))
def test_randomized_shapes_layouts_and_functors(native, case):
    from dcfp_b200 import ops
    i, N, C, h, w, K, layout, dtype, functor, blob = case
    g = torch.Generator().manual_seed(1000 + i)
    scale_up = [1, 2, 4, 8][i % 4]
    H0, W0 = h * scale_up + (i % 3 if scale_up > 1 else 0), w * scale_up + (i % 2 if scale_up > 1 else 0)
    label = _labels(N, H0, W0, K, [torch.uint8, torch.int32, torch.int64][i % 3], seed=2000 + i, blob=blob)
    x = torch.randn(N, C, h, w, generator=g).to(dtype)
    dy = scale = shift = None
    if functor != "fwd":
        scale = torch.rand(C, generator=g) + 0.5
        shift = torch.randn(C, generator=g)
    if functor == "bwd":
        dy = (torch.randn(N, C, h, w, generator=g) * 0.1).to(dtype)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
        if dy is not None:
            dy = dy.contiguous(memory_format=torch.channels_last)
    if dtype == torch.bfloat16:
        # the oracle sees the SAME bf16-rounded inputs widened to fp32; tolerance stays the fp32-accumulation one
        x32, dy32 = x.float(), None if dy is None else dy.float()
    else:
        x32, dy32 = x, dy
    dev = torch.device("cuda")
    S1 = torch.full((K, C + 3), 7.0, dtype=torch.float64, device=dev)  # guard columns: a shared-arena layout, ld = C + 3
    S2 = torch.full_like(S1, 7.0)
    S1[:, :C] = 0
    S2[:, :C] = 0
    cn = torch.zeros(K, dtype=torch.float64, device=dev)
    keys = _keys(ops, label, h, w, K, cn)
    to = lambda t: None if t is None else t.to(dev)
    ops.class_stats(x.to(dev), keys, K, S1[:, :C], S2[:, :C], dy=to(dy), scale=to(scale), shift=to(shift))
    torch.cuda.synchronize()
    v = ref.functor_fwd(x32.float(), scale, shift) if dy is None else ref.functor_bwd(x32.float(), dy32.float(), scale, shift)
    rc, r1, r2 = ref.class_stats(v, label, K)
    mass = ref.abs_mass(v, label, K)
    assert torch.equal(cn.cpu(), rc)
    assert (S1[:, C:] == 7.0).all() and (S2[:, C:] == 7.0).all(), "wrote outside the layer's columns"
    e1 = (S1[:, :C].cpu() - r1).abs()
    assert (e1 <= RTOL * mass + 1e-30).all(), "S1 max rel-to-mass err %.3g" % (e1 / (mass + 1e-30)).max()
    e2 = (S2[:, :C].cpu() - r2).abs()
    assert (e2 <= RTOL * r2 + 1e-30).all(), "S2 max rel err %.3g" % (e2 / (r2 + 1e-30)).max()


@pytest.mark.parametrize("seed", list(range(10)))
def test_randomized_grouped_launches_on_a_shared_arena(native, seed):
    """What the scorer does: many layers of mixed shapes (and mixed layouts) in ONE grouped call, every layer owning a
    column range of one [K, sum C] arena (row stride = sum C), backward functor with autograd-style (invstd, mean).
    Checked per layer against the oracle, plus: nothing outside a layer's columns is touched."""
    import random
    from dcfp_b200 import ops
    rnd = random.Random(900 + seed)
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3000 + seed)
    K = rnd.choice([2, 12, 13, 19, 20, 25, 151, 172])
    bwd = seed % 3 != 0
    dtype = torch.bfloat16 if seed % 5 == 4 else torch.float32
    N = rnd.choice([1, 2, 3])
    res = [(rnd.choice([4, 8, 16, 17, 32]), rnd.choice([4, 8, 16, 24, 33, 64])) for _ in range(3)] + [(1, 1)]
    label = _labels(N, 128, 192, K, torch.uint8, seed=4000 + seed, blob=seed % 2 == 0)
    layers = []
    for _ in range(rnd.randint(3, 14)):
        C = rnd.choice([4, 48, 64, 64, 96, 128, 128, 256, 320, 512, 36, 6])
        h, w = rnd.choice(res)
        nhwc = rnd.random() < 0.6
        layers.append((C, h, w, nhwc))
    total = sum(l[0] for l in layers) + 5
    A1 = torch.full((K, total), 3.0, dtype=torch.float64, device=dev)
    A2 = torch.full_like(A1, 3.0)
    keys = {}
    xs, dys, scs, shs, kl, S1s, S2s, refs = [], [], [], [], [], [], [], []
    pos = 2  # two guard columns in front, three behind
    for C, h, w, nhwc in layers:
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        x = (torch.randn(N, C, h, w, generator=g) * 1.5 + 0.2).to(dtype).contiguous(memory_format=fmt)
        dy = (torch.randn(N, C, h, w, generator=g) * 0.05).to(dtype).contiguous(memory_format=fmt) if bwd else None
        mean = torch.randn(C, generator=g) * 0.3
        invstd = torch.rand(C, generator=g) + 0.5
        if (h, w) not in keys:
            keys[(h, w)] = _keys(ops, label, h, w, K)
        A1[:, pos:pos + C] = 0
        A2[:, pos:pos + C] = 0
        xs.append(x.to(dev)); dys.append(None if dy is None else dy.to(dev)); scs.append(invstd.to(dev)); shs.append(mean.to(dev))
        kl.append(keys[(h, w)]); S1s.append(A1[:, pos:pos + C]); S2s.append(A2[:, pos:pos + C])
        if bwd:
            v = ref.functor_bwd(x.float(), dy.float(), invstd, -mean * invstd)
        else:
            v = ref.functor_fwd(x.float(), invstd, -mean * invstd)
        refs.append((pos, C, v))
        pos += C
    # layers of one call must share dtype and functor; layouts may mix
    ops.class_stats_grouped(xs, kl, K, S1s, S2s, dys=dys if bwd else None, scales=scs, shifts=shs, affine_mode=ops.AFFINE_INVSTD_MEAN)
    torch.cuda.synchronize()
    own = torch.zeros(total, dtype=torch.bool)
    for p, C, v in refs:
        own[p:p + C] = True
        _, r1, r2 = ref.class_stats(v, label, K)
        mass = ref.abs_mass(v, label, K)
        e1 = (A1[:, p:p + C].cpu() - r1).abs()
        # (x - mean) * invstd evaluated as fma(x, invstd, -mean * invstd): one extra rounding of the shift
        assert (e1 <= 2 * RTOL * mass + 1e-30).all(), "columns %d:%d S1 max rel-to-mass err %.3g" % (p, p + C, (e1 / (mass + 1e-30)).max())
        e2 = (A2[:, p:p + C].cpu() - r2).abs()
        assert (e2 <= 4 * RTOL * r2 + 1e-30).all(), "columns %d:%d S2 max rel err %.3g" % (p, p + C, (e2 / (r2 + 1e-30)).max())
    assert (A1.cpu()[:, ~own] == 3.0).all() and (A2.cpu()[:, ~own] == 3.0).all(), "wrote outside the layers' columns"


@pytest.mark.parametrize("warps", ["8", "16"])
def test_forward_kernel_with_a_forced_warp_count(native, warps):
    """The forward functor has an 8-warp and a 16-warp NHWC kernel (default: 16 for bf16 maps, 8 for fp32;
    DCFP_K1_FWD_WARPS, read once per process, forces one for both): the channels_last / randomized / shared-arena cases
    above run again in a child process with each, so both dtypes are checked on both kernels."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, DCFP_K1_FWD_WARPS=warps)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                           "-k", "(channels_last or randomized or shared_arena or maximum_class or more_layers or fwd_matches) "
                                 "and not forced_warp_count"],
                          cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:]
