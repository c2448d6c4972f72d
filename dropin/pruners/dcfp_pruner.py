from dcfp_b200.pruners.dcfp_pruner import DCFPPruner, dcfp_pruning  # noqa: F401
