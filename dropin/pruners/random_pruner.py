from dcfp_b200.pruners.random_pruner import RandomChannelPruner  # noqa: F401
