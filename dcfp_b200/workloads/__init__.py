from .segnets import CONFIGS, BACKBONE_PARA, SegNet, CalibrationLoss, build_segnet  # noqa: F401
from .synthetic import synthetic_images, synthetic_labels, class_prior  # noqa: F401
