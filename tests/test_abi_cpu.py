"""The C-ABI shared library: loads without a GPU, exports every symbol include/dcfp_b200.h declares,
and rejects bad arguments with a negative code and a message (validation happens before any CUDA call)."""
import ctypes
import os
import subprocess
import sys

import pytest

from dcfp_b200 import abi, build


@pytest.fixture(scope="module")
def lib():
    build.build_abi()
    return abi.load()


def test_exports_every_declared_symbol(lib):
    names = abi.declared_symbols()
    assert len(names) >= 15 and "dcfp_class_stats_grouped" in names and "dcfp_fold_step" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    leaked = [s for s in exported if not s.startswith("dcfp_") and not s.startswith("_")]
    assert not leaked, "non-ABI symbols exported: %s" % leaked[:5]


def test_abi_version_and_struct_layout(lib):
    assert lib.dcfp_abi_version() == 4
    assert ctypes.sizeof(abi.LayerDesc) == 7 * 8 + 10 * 4
    assert ctypes.sizeof(abi.GatherDesc) == 4 * 8 + 4 * 4
    assert ctypes.sizeof(abi.BnDesc) == 16 * 8 + 8 * 4 + 2 * 4 + 2 * 4 + 3 * 8  # struct dcfp_bn_desc
    assert lib.dcfp_bn_scratch_bytes(256) >= 4 * 2 * 256 * 8 + 5 * 256 * 4 and lib.dcfp_bn_workspace_bytes(256) > 0
    assert lib.dcfp_bn_supported(2, 256, 64, 128, abi.F32) == 1 and lib.dcfp_bn_supported(2, 30, 64, 128, abi.F32) == 0
    assert lib.dcfp_channel_gather_workspace(10) >= 10 * ctypes.sizeof(abi.GatherDesc) + 11 * 8


def test_validation_errors_do_not_touch_the_device(lib):
    d = abi.LayerDesc()
    assert lib.dcfp_class_stats(ctypes.byref(d), None) == -1  # DCFP_EINVAL: null x / S1 / S2
    assert b"non-null" in lib.dcfp_last_error()
    d.x = d.S1 = d.S2 = 0x1000
    d.N, d.C, d.h, d.w, d.K = 2, 8, 4, 4, 300
    assert lib.dcfp_class_stats(ctypes.byref(d), None) == -3  # DCFP_ETOOBIG: K > 255
    assert b"K=300" in lib.dcfp_last_error()
    d.K, d.keys = 19, 0x1000
    d.affine_mode = 1  # INVSTD_MEAN without scale / shift
    assert lib.dcfp_class_stats(ctypes.byref(d), None) == -1
    assert lib.dcfp_label_keys(None, 0, 1, 1, 1, 1, 1, 19, None, None, None) == -1
    assert lib.dcfp_eic_update_flat(None, None, None, 4, 0.5, 0.5, 1, None) == -1
    assert lib.dcfp_fold_step(None, None, 19, 4, None, None) == -1
    assert lib.dcfp_fold_step2(None, None, None, 19, 4, None, None) == -1
    k = (ctypes.c_int64 * 2)(5, 0)
    assert lib.dcfp_thresh_mask(0x1000, 0x1000, 0x1000, 0x1000, 1, 4, k, 0x1000, 0x1000, None, None) == -1  # k_idx >= n_total
    assert lib.dcfp_channel_gather(0x1000, 0x1000, None, 4, None, 3, 4, 1, 4, None) == -1  # in_idx NULL but n_in != I
    assert lib.dcfp_channel_gather(0x1000, 0x1000, None, 4, None, 4, 4, 1, 8, None) == -2  # DCFP_EUNSUPPORTED elt size
    assert lib.dcfp_bias_comp(None, 1, 1, 1, None, None, None) == -1
    b = abi.BnDesc()
    assert lib.dcfp_bn_forward(ctypes.byref(b), None) == -1 and b"null pointer" in lib.dcfp_last_error()
    assert lib.dcfp_bn_backward(None, None) == -1


def test_product_path_fails_loudly_without_gpu():
    """No CPU fallback: product entry points raise when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dcfp_b200 import ops
    from dcfp_b200.pruners.dcfp_pruner import dcfp_pruning
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.require_gpu()
    bn = torch.nn.Sequential(torch.nn.BatchNorm2d(4))
    bn.ignore_prune_layer = []
    tp = dcfp_pruning(bn, 0.9)
    bn[0].weight.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="GPU only"):
        tp.step(bn)
    # the registered torch ops have no CPU kernels either
    ops.load()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.dcfp.reduce_classes(torch.zeros(2, 3, dtype=torch.float64))


def test_no_product_module_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dirpath, _, files in os.walk(os.path.join(root, "dcfp_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                if "import oracle" in text or "from oracle" in text:
                    bad.append(f)
    assert not bad, bad
    code = "import sys; import dcfp_b200.scorer, dcfp_b200.pruners.dcfp_pruner, dcfp_b200.pruners.random_pruner; " \
           "assert not [m for m in sys.modules if m.split('.')[0] == 'oracle']"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root)


def test_python_constants_match_the_kernel_headers():
    """ops.K1_FWD_WARPS_BF16 mirrors kNhwcFwdWarpsBf16."""
    import re

    from dcfp_b200 import ops

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "dcfp_b200", "csrc", "k1_nhwc.cuh")).read()
    m = re.search(r"constexpr int kNhwcFwdWarpsBf16 = (\d+);", src)
    assert m and int(m.group(1)) == ops.K1_FWD_WARPS_BF16
