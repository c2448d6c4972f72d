"""TEST-ONLY launcher: run one of the reference's own CLIs (prune.py) UNMODIFIED, with `import pruners`
resolving either to the reference's package or to this repo's drop-in (dropin/pruners -> dcfp_b200.pruners).

    python tests/run_reference_cli.py {reference|dropin} [--oracle-backend] /root/reference/prune.py <prune.py args>

Mechanics: the script directory is NOT put on sys.path (the effect of `python -P`), the
drop-in directory is placed before the reference tree, torch.load is wrapped with weights_only=False
(prune.py:108 loads numpy arrays).  `--oracle-backend` swaps the CUDA kernels for the CPU oracle so the HOST
side of the drop-in can be exercised in a container without a GPU; it is never used by the product."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    which, args = sys.argv[1], sys.argv[2:]
    oracle = "--oracle-backend" in args
    if oracle:
        args.remove("--oracle-backend")
    script, args = args[0], args[1:]
    ref_root = os.path.dirname(os.path.abspath(script))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ref_root, os.path.join(ROOT, "tests"))]
    sys.path.insert(0, ROOT)
    from oracle import ref_compat
    ref_compat.REF_ROOT = ref_root
    if which == "reference":
        ref_compat.load_reference()  # reference pruners + the arithmetic-free shims
    else:
        import types
        shim = types.ModuleType("ordered_set")
        shim.OrderedSet = ref_compat.OrderedSet
        sys.modules.setdefault("ordered_set", shim)
        sys.path.insert(0, os.path.join(ROOT, "dropin"))
        sys.path.append(ref_root)
        import pruners
        assert pruners.__file__.startswith(os.path.join(ROOT, "dropin")), pruners.__file__
    sys.setrecursionlimit(100000)
    sys.argv = [script] + args
    with ref_compat.legacy_torch_load():
        if oracle and which == "dropin":
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from fake_backend import oracle_backend
            with oracle_backend():
                runpy.run_path(script, run_name="__main__")
        else:
            runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
