"""oracle/soft_argmax_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Op-for-op torch restatement of the reference's hot path, runnable on CPU:

* ``ref_soft_argmax``      follows /root/reference/common/nets/loss.py:13-34
* ``RefJointLocationLoss`` follows /root/reference/common/nets/loss.py:36-52

The only deviations from the reference text are the ones that make it runnable off-GPU and
shape-generic: ``cfg.depth_dim`` / ``cfg.output_shape`` (loss.py:16,18) are inferred from the
tensor instead of read from a global, and the index weights ``arange(1, n+1)`` are created on
the input's device instead of going through ``torch.cuda.FloatTensor`` +
``torch.cuda.comm.broadcast`` (loss.py:24-26).  It issues the same ATen calls in the same
order, so on the same machine it is bit-identical to the unmodified reference run through the
two-attribute shim (checked by oracle/make_golden.py when the goldens are produced, and by
tests/test_oracle.py against the committed goldens).

Pinned: yes -- against tests/golden/*.npz produced by executing the reference itself.
This is also the "port" that bench.py times as the CPU baseline (it is what the reference
executes on a CPU: torch eager ops, all host threads).
"""
import torch
from torch.nn import functional as F


def ref_soft_argmax(heatmaps, joint_num, depth_dim=None):
    assert isinstance(heatmaps, torch.Tensor)                                   # loss.py:14
    H, W = heatmaps.shape[-2], heatmaps.shape[-1]
    D = depth_dim if depth_dim is not None else heatmaps.shape[1] // joint_num
    heatmaps = heatmaps.reshape((-1, joint_num, D * H * W))                     # loss.py:16
    heatmaps = F.softmax(heatmaps, 2)                                           # loss.py:17
    heatmaps = heatmaps.reshape((-1, joint_num, D, H, W))                       # loss.py:18

    accu_x = heatmaps.sum(dim=(2, 3))                                           # loss.py:20
    accu_y = heatmaps.sum(dim=(2, 4))                                           # loss.py:21
    accu_z = heatmaps.sum(dim=(3, 4))                                           # loss.py:22

    dev = heatmaps.device
    accu_x = accu_x * torch.arange(1, W + 1).type(torch.FloatTensor).to(dev)    # loss.py:24
    accu_y = accu_y * torch.arange(1, H + 1).type(torch.FloatTensor).to(dev)    # loss.py:25
    accu_z = accu_z * torch.arange(1, D + 1).type(torch.FloatTensor).to(dev)    # loss.py:26

    accu_x = accu_x.sum(dim=2, keepdim=True) - 1                                # loss.py:28
    accu_y = accu_y.sum(dim=2, keepdim=True) - 1                                # loss.py:29
    accu_z = accu_z.sum(dim=2, keepdim=True) - 1                                # loss.py:30

    return torch.cat((accu_x, accu_y, accu_z), dim=2)                           # loss.py:32


def _assert_no_grad(tensor):                                                    # loss.py:8-11
    assert not tensor.requires_grad, \
        "nn criterions don't compute the gradient w.r.t. targets - please " \
        "mark these tensors as not requiring gradients"


class RefJointLocationLoss(torch.nn.Module):
    def forward(self, heatmap_out, gt_coord, gt_vis, gt_have_depth):
        joint_num = gt_coord.shape[1]                                           # loss.py:42
        coord_out = ref_soft_argmax(heatmap_out, joint_num)                     # loss.py:43
        _assert_no_grad(gt_coord)                                               # loss.py:45-47
        _assert_no_grad(gt_vis)
        _assert_no_grad(gt_have_depth)
        loss = torch.abs(coord_out - gt_coord) * gt_vis                         # loss.py:49
        loss = (loss[:, :, 0] + loss[:, :, 1] + loss[:, :, 2] * gt_have_depth) / 3.   # loss.py:50
        return loss.mean()                                                      # loss.py:52


def ref_final_layer(x, weight, bias):
    """HeadNet.final_layer: Conv2d(256 -> J*D, k=1) with bias.  /root/reference/main/model.py:14-20,42"""
    return F.conv2d(x, weight.reshape(weight.shape[0], weight.shape[1], 1, 1), bias)


def ref_fwd_bwd(heat, gt, vis, have_depth):
    """One reference training 'step' of the path: loss forward + autograd backward.
    Returns (loss, coords, grad_heat).  Mirrors /root/reference/main/train.py:67-71."""
    h = heat.detach().clone().requires_grad_(True)
    coords = ref_soft_argmax(h, gt.shape[1])
    loss = torch.abs(coords - gt) * vis
    loss = ((loss[:, :, 0] + loss[:, :, 1] + loss[:, :, 2] * have_depth) / 3.).mean()
    loss.backward()
    return loss.detach(), coords.detach(), h.grad
