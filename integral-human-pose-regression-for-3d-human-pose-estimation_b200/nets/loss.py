"""Drop-in for /root/reference/common/nets/loss.py -- same names, arguments and error behaviour.

``soft_argmax`` (loss.py:13-34) and ``JointLocationLoss`` (loss.py:36-52) dispatch to the sm_100a
kernels; ``JointMSELoss`` (loss.py:55-84, the direct-regression baseline's criterion, off the hot
path) is plain torch so ``common/base.py:207`` keeps importing.

Differences a caller can observe: the volume shape is inferred from the tensor (C // joint_num,
H, W) instead of read from the global ``cfg`` (loss.py:16,18) -- when a ``config`` module with
``cfg`` is loaded it is cross-checked; CPU tensors raise (the reference also only runs on CUDA,
loss.py:24-26).
"""
import sys

import torch
import torch.nn as nn

from ..functional import DeferredHeatmap, integral_l1_loss, integral_l1_step
from ..functional import soft_argmax as _soft_argmax


def _assert_no_grad(tensor):                                             # loss.py:8-11
    assert not tensor.requires_grad, \
        "nn criterions don't compute the gradient w.r.t. targets - please " \
        "mark these tensors as not requiring gradients"


def _check_cfg(heatmaps, joint_num):
    cfg_mod = sys.modules.get("config")
    cfg = getattr(cfg_mod, "cfg", None)
    if cfg is None or not hasattr(cfg, "depth_dim") or not hasattr(cfg, "output_shape"):
        return
    C, H, W = heatmaps.shape[1], heatmaps.shape[2], heatmaps.shape[3]
    if C != joint_num * cfg.depth_dim or (H, W) != tuple(cfg.output_shape):
        raise ValueError("heatmaps %s disagree with cfg (depth_dim=%s, output_shape=%s, joint_num=%d)"
                         % (tuple(heatmaps.shape), cfg.depth_dim, cfg.output_shape, joint_num))


def soft_argmax(heatmaps, joint_num):
    assert isinstance(heatmaps, (torch.Tensor, DeferredHeatmap))         # loss.py:14
    _check_cfg(heatmaps, joint_num)
    return _soft_argmax(heatmaps, joint_num)


class JointLocationLoss(nn.Module):
    def __init__(self, fused_backward=None):
        """fused_backward: None = library default (on), see functional.integral_l1_loss."""
        super(JointLocationLoss, self).__init__()
        self.fused_backward = fused_backward

    def forward(self, heatmap_out, gt_coord, gt_vis, gt_have_depth):
        joint_num = gt_coord.shape[1]                                    # loss.py:42
        _assert_no_grad(gt_coord)                                        # loss.py:45-47
        _assert_no_grad(gt_vis)
        _assert_no_grad(gt_have_depth)
        _check_cfg(heatmap_out, joint_num)
        return integral_l1_loss(heatmap_out, gt_coord, gt_vis, gt_have_depth, fused_backward=self.fused_backward)

    def forward_backward(self, heatmap_out, gt_coord, gt_vis, gt_have_depth):
        """``loss = criterion(...); loss.backward()`` (main/train.py:67-71) as one call and ONE launch: the gradient K5 produced with
        the loss goes straight into autograd (no ones-fill, no rescale launch).  Returns the detached loss."""
        _assert_no_grad(gt_coord)
        _assert_no_grad(gt_vis)
        _assert_no_grad(gt_have_depth)
        _check_cfg(heatmap_out, gt_coord.shape[1])
        return integral_l1_step(heatmap_out, gt_coord, gt_vis, gt_have_depth)[0]


class JointMSELoss(nn.Module):
    """Criterion of the direct-regression baseline (loss.py:55-84; used by common/base.py:207), off the accelerated
    path.  Same value as the reference's per-joint loop -- sum_j 0.5 * mean_{b,c}((w_bj * (pred_bjc - gt_bjc))^2) / J --
    computed in one vectorised expression."""

    def __init__(self):
        super(JointMSELoss, self).__init__()
        self.use_target_weight = True

    def forward(self, output, target, target_weight):
        _assert_no_grad(target)
        _assert_no_grad(target_weight)
        batch, joints = target.size(0), target.size(1)
        diff = output.reshape(batch, joints, -1) - target.reshape(batch, joints, -1)
        if self.use_target_weight:
            diff = diff * target_weight.reshape(batch, joints, 1)
        return 0.5 * diff.pow(2).mean(dim=(0, 2)).sum() / joints
