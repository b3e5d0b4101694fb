"""nopesac_b200 — B200-native (sm_100a) implementation of NopeSAC's one-plane RANSAC pose path
(CAMERA_HEAD hypothesis loop + the plane matching that feeds it) behind the reference's registry surface.

The CUDA kernels live in csrc/ and are reached through the C ABI of include/nopesac_b200.h
(libnopesac_b200.so, loaded with ctypes).  There is no CPU or eager-PyTorch fallback.
"""
from .backbone import BACKBONE_REGISTRY, ResNet50Backbone, build_backbone, build_resnet_backbone  # noqa: F401
from .camera_head import CAMERA_HEAD_REGISTRY, PlaneCameraHead, build_camera_head  # noqa: F401
from .config import get_cfg, get_sparseplane_cfg_defaults, load_config  # noqa: F401
from .matching_head import MATCHING_HEAD_REGISTRY, MatchingHead, build_matching_head  # noqa: F401
from .meta_arch import META_ARCH_REGISTRY, PlaneTR_NopeSAC, build_model  # noqa: F401

__version__ = "0.1.0"
