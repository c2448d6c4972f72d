"""Class-balance pixel weights (datasets/Base.py:73-89) on the GPU vs the reference's own outputs (tests/golden/balance.npz)
and the CPU oracle: counts exact, weights to 1e-12 relative (CUDA's double pow is not bit-identical to numpy's)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import balance_ref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_class_balance_weights_vs_reference_golden(native):
    from dcfp_b200 import ops
    z = np.load(os.path.join(GOLDEN, "balance.npz"))
    cases = json.loads(bytes(z["meta"]).decode())
    by_case = {}
    for c in cases:
        by_case.setdefault(c["case"], []).append(c)
    for ci, group in by_case.items():  # one batched call per case: N images at once
        labels = np.stack([z["label_%d_%d" % (ci, c["img"])] for c in group])
        cls = torch.tensor([c["sample_class"] for c in group], dtype=torch.int32, device="cuda")
        for dtype in (torch.uint8, torch.int64):
            w, cnt = ops.class_balance_weights(torch.from_numpy(labels).to("cuda").to(dtype), group[0]["K"], cls,
                                               mode=group[0]["balance"], beta=group[0]["beta"])
            w, cnt = w.cpu().numpy(), cnt.cpu().numpy()
            for i, c in enumerate(group):
                exp = z["weight_%d_%d" % (ci, c["img"])]
                assert w[i].dtype == np.float64 and np.allclose(w[i], exp, rtol=1e-12, atol=0), c
                _, cnt_ref = balance_ref.class_balance_weights(labels[i], c["K"], c["sample_class"], c["balance"], c["beta"])
                assert np.array_equal(cnt[i], cnt_ref)
                assert (w[i][labels[i] == 255] == 0).all() and w[i].max() <= 1.0


def test_class_balance_validation(native):
    from dcfp_b200 import ops
    lab = torch.zeros(2, 8, 8, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="mode 2 needs sample_class"):
        ops.class_balance_weights(lab, 19, None, mode=2)
    with pytest.raises(RuntimeError, match="mode 3"):
        ops.class_balance_weights(lab, 19, None, mode=3)
    w, cnt = ops.class_balance_weights(lab, 19, None, mode=1)
    assert cnt[:, 0].tolist() == [64, 64] and torch.allclose(w, torch.full_like(w, 1.0 / 65))
