"""baseline/_ref/ (what travels to the GPU box) is a byte-identical copy of the reference's Python sources."""
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.isfile(os.path.join(DST, "MANIFEST.json")), reason="baseline/_ref not vendored (scripts/vendor_reference.py)")
def test_vendored_reference_is_unmodified():
    with open(os.path.join(DST, "MANIFEST.json")) as f:
        man = json.load(f)
    assert "pruners/dcfp_pruner.py" in man["files"] and "prune.py" in man["files"]
    for rel, digest in man["files"].items():
        with open(os.path.join(DST, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference", rel)
        if os.path.isfile(src):  # build container: compare with the read-only tree itself
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel
