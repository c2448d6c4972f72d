import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the unmodified reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch

    from oracle import ref_compat

    has_gpu = torch.cuda.is_available()
    has_ref = ref_compat.available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "ref" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present (GPU box)"))


@pytest.fixture(scope="session")
def native():
    """Builds (if stale) and loads the native library; GPU tests call through torch.ops.dcfp -> C ABI."""
    from dcfp_b200 import build, ops

    build.build_all()
    return ops.load()
