from dcfp_b200.pruners.channel_pruner import *  # noqa: F401,F403
from dcfp_b200.pruners.channel_pruner import ChannelPruner, init_pruned_model  # noqa: F401
