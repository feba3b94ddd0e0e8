"""ihpr_b200 -- B200-native (sm_100a) integral-regression hot path.

Drop-in for the reference's ``common/nets/loss.py`` callables (``soft_argmax``,
``JointLocationLoss``) backed by hand-written CUDA kernels behind a C-ABI library
(``include/ihpr_b200.h``).  There is no CPU fallback: without the built library or without an
sm_100 device every op raises.
"""
from .functional import (soft_argmax, integral_l1_loss, integral_l1_step, integral_l1_fwd_bwd_host, last_launch_count, last_path_choice,  # noqa: F401
                         set_variant, get_variant, fused_head_soft_argmax, fused_head_integral_l1_loss, flip_merge, flip_perm, DeferredHeatmap,
                         coords_to_camera, deconv_bn_relu, deconv_bn_relu_train)
from .nets.loss import JointLocationLoss, JointMSELoss  # noqa: F401
from .dropin import install_dropin  # noqa: F401
from .data import augment_batch, get_aug_config, gen_trans_from_patch  # noqa: F401
from ._lib import IhprError, library_path, version  # noqa: F401

__all__ = ["soft_argmax", "integral_l1_loss", "integral_l1_step", "integral_l1_fwd_bwd_host", "JointLocationLoss", "JointMSELoss",
           "install_dropin", "fused_head_soft_argmax", "fused_head_integral_l1_loss", "flip_merge", "flip_perm", "DeferredHeatmap", "coords_to_camera", "deconv_bn_relu", "deconv_bn_relu_train", "augment_batch", "get_aug_config", "gen_trans_from_patch", "IhprError", "library_path", "version", "set_variant", "get_variant",
           "last_launch_count", "last_path_choice"]
