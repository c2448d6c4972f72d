"""Drop-in `pruners` package for the reference's own CLIs (prune.py:11-13, train.py:22, evaluate.py:20).

    PYTHONPATH=<repo>/dropin:<repo>:<reference> python -P <reference>/prune.py ...

`-P` keeps the script directory off sys.path so `import pruners` resolves here while `networks`,
`utils` still resolve to the reference tree.
"""
from dcfp_b200.pruners.channel_pruner import init_pruned_model  # noqa: F401
from dcfp_b200.pruners.dcfp_pruner import dcfp_pruning  # noqa: F401
