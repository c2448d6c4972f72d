from .channel_pruner import init_pruned_model  # noqa: F401
from .dcfp_pruner import dcfp_pruning  # noqa: F401
