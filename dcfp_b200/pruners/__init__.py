"""Host-side mirror of the reference's `pruners` package (same public names; see dropin/pruners for the drop-in).

The two names the reference's own `pruners/__init__.py` exports are what `train.py` / `evaluate.py` reach through
`import pruners` (train.py:202,216; evaluate.py:289); the pruner classes live in the sub-modules, as in the reference.
"""
from .dcfp_pruner import dcfp_pruning
from .channel_pruner import init_pruned_model

__all__ = ["dcfp_pruning", "init_pruned_model"]
