"""The reference's UNMODIFIED prune.py (prune.py:91-124) on the GPU box: once with its own `pruners` package, once with
`dropin/pruners` on the CUDA backend (K2 thresholds / masks, K3 gather, bias compensation kernel -- no oracle anywhere in
that process).  Same global_percent trajectory; channel_cfg.pth and pruned.pth bit-identical.

The reference tree is not on the GPU box: scripts/vendor_reference.py leaves a byte-identical copy of its Python sources in
the git-ignored baseline/_ref/ (which gpurun ships); the test is skipped when no reference tree can be found."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import ref_compat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.ref]


def test_unmodified_prune_py_cuda_backend_equals_reference_pruners(native, tmp_path):
    model = gu.build_model("c1")
    ckpt, score = str(tmp_path / "model.pth"), str(tmp_path / "score.pth")
    torch.save(dict(model.state_dict()), ckpt)
    eic = gu.make_scores(model, "uniform", 31)
    torch.save({"eic": {k: torch.from_numpy(v) for k, v in eic.items()}}, score)
    procs = {}
    for which in ("reference", "dropin"):  # the two CLIs run side by side (the FLOPs counter's CPU forwards dominate)
        save = str(tmp_path / which)
        cmd = [sys.executable, os.path.join(ROOT, "tests", "run_reference_cli.py"), which,
               os.path.join(ref_compat.REF_ROOT, "prune.py"), "--model", "deeplabv3", "--backbone", "resnet50",
               "--backbone-para", '{"os": 8, "mg_unit": [1,2,4], "inplanes": 128, "pretrained": false}',
               "--dataset", "CS", "--prune-ratio", "0.45", "--model-path", ckpt, "--score-path", score, "--save-path", save]
        env = dict(os.environ, DCFP_TRACE_BACKEND="1")
        procs[which] = (save, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), env=env))
    outs = {}
    for which, (save, p) in procs.items():
        so, se = p.communicate(timeout=900)
        assert p.returncode == 0, so[-2000:] + se[-2000:]
        lines = [l for l in so.splitlines() if l.startswith(("global_percent", "flops", "Finish"))]
        outs[which] = (lines, torch.load(os.path.join(save, "pruned.pth"), weights_only=False),
                       torch.load(os.path.join(save, "channel_cfg.pth"), weights_only=False), so)
    assert "dcfp backend: cuda" in outs["dropin"][3], "the drop-in process did not run on the CUDA kernels"
    assert outs["reference"][0] == outs["dropin"][0] and any(l.startswith("Finish") for l in outs["dropin"][0])
    a, b = outs["reference"][1], outs["dropin"][1]
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k  # random-init BN beta = 0: no bias compensation, every tensor bit-exact
    ca, cb = outs["reference"][2], outs["dropin"][2]
    assert list(ca.keys()) == list(cb.keys())
    for k in ca:
        for kk, v in ca[k].items():
            assert np.array_equal(v, cb[k][kk]) if isinstance(v, np.ndarray) else v == cb[k][kk], (k, kk)
