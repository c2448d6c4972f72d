"""Import the UNMODIFIED reference (wzx99/DCFP) in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `/root/reference` does not exist
on the GPU box; ``scripts/vendor_reference.py`` leaves a byte-identical copy of its Python
sources in the git-ignored ``baseline/_ref/``, which does travel.  Used (a) by
``tests/golden/make_golden.py`` to generate committed fixtures, (b) by tests marked ``ref``
(skipped when no reference tree is found) and (c) by ``bench.py --impl reference``.

Three external shims are needed to run the reference under torch 2.11 / py3.12
(SURVEY.md section 8c); none of them carries arithmetic:

1. ``ordered_set`` (third-party, un-vendored, unpinned -- reference README.md:18) is
   not installed: a minimal insertion-ordered set stands in
   (used by pruners/channel_pruner.py:9,246,443).
2. torch >= 1.11 names the conv autograd node ``ConvolutionBackward0`` which is absent
   from ``CONV`` (pruners/channel_pruner.py:12-13,621-624): register it.
3. ``torch.load`` defaults to ``weights_only=True`` since torch 2.6, which rejects the
   numpy arrays inside ``channel_cfg.pth`` (prune.py:108): wrap with
   ``weights_only=False``.
"""
import os
import sys
import types
import contextlib

def _find_reference():
    """$DCFP_REF, else the read-only tree of the build container, else the byte-identical copy that
    scripts/vendor_reference.py leaves in the git-ignored baseline/_ref/ (the only one present on the GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("DCFP_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "pruners", "dcfp_pruner.py")):
            return cand
    return os.environ.get("DCFP_REF", "/root/reference")


REF_ROOT = _find_reference()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "pruners", "dcfp_pruner.py"))


class OrderedSet:
    """Insertion-ordered set with the handful of methods the reference touches."""

    def __init__(self, items=()):
        self._d = dict.fromkeys(items)

    def add(self, x):
        self._d.setdefault(x, None)

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, x):
        return x in self._d

    def __getitem__(self, i):
        return list(self._d)[i]

    def intersection(self, other):
        o = set(other)
        return OrderedSet(x for x in self._d if x in o)

    def union(self, other):
        return OrderedSet(list(self._d) + list(other))

    def __repr__(self):
        return "OrderedSet(%r)" % (list(self._d),)


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's modules (pruners, networks, loss)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not found at %s (set DCFP_REF)" % REF_ROOT)
    if "ordered_set" not in sys.modules:
        shim = types.ModuleType("ordered_set")
        shim.OrderedSet = OrderedSet
        sys.modules["ordered_set"] = shim
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    clash = sys.modules.get("pruners")
    if clash is not None and not getattr(clash, "__file__", "").startswith(REF_ROOT):
        raise RuntimeError("a non-reference top-level `pruners` is already imported")
    import pruners  # noqa: the reference's
    import pruners.channel_pruner as cp
    import pruners.dcfp_pruner as dp
    import pruners.random_pruner as rp
    import networks
    import loss.criterion as crit

    assert cp.__file__.startswith(REF_ROOT), cp.__file__
    if "ConvolutionBackward" not in cp.CONV:
        cp.CONV = cp.CONV + ("ConvolutionBackward",)
        cp.NON_PASS = cp.CONV + cp.FC
        cp.BACKWARD_PARSER_DICT["ConvolutionBackward"] = cp.ChannelPruner.conv_backward_parser
    _loaded.update(pruners=pruners, cp=cp, dp=dp, rp=rp, networks=networks, crit=crit)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def legacy_torch_load():
    import torch

    orig = torch.load

    def _load(*a, **kw):
        kw.setdefault("weights_only", False)
        return orig(*a, **kw)

    torch.load = _load
    try:
        yield
    finally:
        torch.load = orig
