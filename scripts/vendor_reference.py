"""Copy the UNMODIFIED reference tree into the git-ignored `baseline/_ref/` so that it travels to the GPU box with
`gpurun` (which ships the repo snapshot, not /root/reference).  Nothing is patched: the files are byte-identical copies
(a manifest with their SHA-256 is written next to them and checked by tests/test_reference_vendor_cpu.py).  Only the Python
sources the hot path's callers need are taken -- the 16 MB of dataset list files stay behind.

    python scripts/vendor_reference.py [/root/reference]

`baseline/_ref/` is test / baseline infrastructure: bench.py --impl reference times the reference's own modules from it
(kind "reference"), tests/test_gpu_dropin_reference.py runs the reference's unmodified prune.py against dropin/pruners on
the CUDA backend and against the reference's own pruners, and compares the outputs bit for bit.  It is listed in
.gitignore (no reference sources in the history) and NOT in .gpurunignore.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_DIRS = ("pruners", "networks", "loss", "utils")
KEEP_FILES = ("prune.py", "train.py", "engine.py", "mypath.py", "optimizer.py", "README.md")


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main(src):
    if not os.path.isfile(os.path.join(src, "pruners", "dcfp_pruner.py")):
        raise SystemExit("no reference tree at %s" % src)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for d in KEEP_DIRS:
        for base, _, files in os.walk(os.path.join(src, d)):
            for fn in files:
                if fn.endswith((".py", ".md")):
                    p = os.path.join(base, fn)
                    rel = os.path.relpath(p, src)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(p, os.path.join(DST, rel))
                    manifest[rel] = sha(p)
    for fn in KEEP_FILES:
        p = os.path.join(src, fn)
        if os.path.isfile(p):
            shutil.copyfile(p, os.path.join(DST, fn))
            manifest[fn] = sha(p)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    print("vendored %d files (%d bytes) into %s" % (len(manifest), sum(os.path.getsize(os.path.join(DST, r)) for r in manifest), DST))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
