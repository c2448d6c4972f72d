"""dcfp_b200 -- B200-native DCFP filter-importance scoring -> keep-mask -> channel-gather path.

Layout:
  csrc/         hand-written sm_100a kernels + the C ABI (include/dcfp_b200.h) + TORCH_LIBRARY binding
  ops.py        loader + thin wrappers over torch.ops.dcfp.* (raises when the CUDA library is missing)
  abi.py        ctypes view of the same C ABI (symbol checks, non-torch callers)
  pruners/      host-side mirror of the reference's pruner API (dcfp_pruning, DCFPPruner, ChannelPruner, ...)
  scorer.py     calibration scorer: BN hooks -> class-stats arena -> (NCCL all-reduce) -> scores
  workloads/    synthetic calibration inputs and the segmentation nets that produce the feature maps
"""
__version__ = "0.1.0"
