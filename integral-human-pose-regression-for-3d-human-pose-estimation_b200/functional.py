"""Host-side operators of the hot path: torch tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, the current stream, autograd bookkeeping); all arithmetic
runs in lib/libihpr_b200.so.  Semantics follow /root/reference/common/nets/loss.py:13-52.
"""
import os
import threading

import weakref

import torch

from . import _lib
from ._lib import IHPR_BF16, IHPR_F32, IhprError, check, lib

_ws_lock = threading.Lock()
_ws_cache = {}      # (device index, stream handle, B*J) -> zero-initialised uint8 workspace


def _dtype_code(t):
    if t.dtype == torch.float32:
        return IHPR_F32
    if t.dtype == torch.bfloat16:
        return IHPR_BF16
    raise TypeError("ihpr_b200 kernels take float32 or bfloat16 heatmaps, got %s" % t.dtype)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise IhprError("ihpr_b200: %s is on %s; this path runs on a CUDA sm_100 device only "
                        "(there is no CPU fallback)" % (name, t.device))


def _shape(heat, joint_num):
    if heat.dim() != 4:
        raise ValueError("heatmaps must be (B, J*D, H, W), got %s" % (tuple(heat.shape),))
    B, C, H, W = heat.shape
    if joint_num <= 0 or C % joint_num != 0:
        raise ValueError("channel count %d is not a multiple of joint_num %d" % (C, joint_num))
    return B, C // joint_num, H, W


def _workspace(dev, stream, nbytes, rows):
    # the kernels leave the workspace's tickets zeroed, but WHERE the tickets live depends on B*J
    # (include/ihpr_b200.h): one workspace per (device, stream, B*J)
    key = (dev.index, stream, rows)
    with _ws_lock:
        ws = _ws_cache.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(max(nbytes, 1), dtype=torch.uint8, device=dev)
            _ws_cache[key] = ws
    return ws


def _f32(t, dev, shape, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.device != dev:
        raise IhprError("%s is on %s but the heatmaps are on %s" % (name, t.device, dev))
    t = t.detach()
    if t.numel() != shape[0] * shape[1] * (shape[2] if len(shape) > 2 else 1):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), shape))
    return t.to(torch.float32).contiguous()


_raw_stream = torch._C._cuda_getCurrentRawStream       # (device index) -> cudaStream_t of torch's current stream, no Stream object
_ws_plan = {}       # (device index, stream, B, J, D, H, W) -> (workspace tensor, data_ptr, nbytes): the steady-state call does one dict lookup


class _on_device:
    """`with torch.cuda.device(dev)` without its cost when `dev` already is the current device (the usual case)."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return _raw_stream(self.idx)

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _planned_workspace(dev, stream, B, J, D, H, W):
    key = (dev.index, stream, B, J, D, H, W)
    plan = _ws_plan.get(key)
    if plan is None:
        nbytes = lib().ihpr_workspace_bytes(B, J, D, H, W)
        ws = _workspace(dev, stream, nbytes, B * J)
        plan = _ws_plan[key] = (ws, ws.data_ptr(), ws.numel())
    return plan


def _fwd(heat, joint_num, targets=None, want_stats=True):
    """heat: contiguous cuda f32/bf16.  Returns (coords, stats, loss-or-None).  One allocation for all the small outputs."""
    B, D, H, W = _shape(heat, joint_num)
    dev = heat.device
    L = lib()
    R = B * joint_num
    with _on_device(dev) as stream:
        _, ws_ptr, ws_n = _planned_workspace(dev, stream, B, joint_num, D, H, W)
        small = torch.empty(R * 5 + 1, dtype=torch.float32, device=dev)
        base = small.data_ptr()
        coords = small[:R * 3].view(B, joint_num, 3)
        stats = small[R * 3:R * 5].view(B, joint_num, 2) if want_stats else None
        if targets is None:
            check(L.ihpr_softargmax3d_fwd(heat.data_ptr(), _dtype_code(heat), B, joint_num, D, H, W,
                                          base, base + R * 12 if want_stats else None, ws_ptr, ws_n, stream))
            return coords, stats, None
        gt, vis, hd = targets
        loss = small[R * 5:].view(())
        check(L.ihpr_integral_l1_fwd(heat.data_ptr(), _dtype_code(heat), B, joint_num, D, H, W,
                                     gt.data_ptr(), vis.data_ptr(), hd.data_ptr(), base + R * 20,
                                     base, base + R * 12, ws_ptr, ws_n, stream))
        return coords, stats, loss


class _SoftArgmax3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heat, joint_num):
        heat = heat.contiguous()
        coords, stats, _ = _fwd(heat, joint_num)
        ctx.joint_num = joint_num
        ctx.variant = lib().ihpr_get_variant()          # the variant is per thread; backward runs on autograd's thread
        ctx.save_for_backward(heat, coords, stats)      # heat is the conv output autograd keeps anyway; no softmax saved
        return coords

    @staticmethod
    def backward(ctx, grad_coords):
        heat, coords, stats = ctx.saved_tensors
        J = ctx.joint_num
        B, D, H, W = _shape(heat, J)
        g = grad_coords.to(torch.float32).contiguous()
        grad_heat = torch.empty_like(heat)
        with _on_device(heat.device) as stream:
            lib().ihpr_set_variant(ctx.variant)
            check(lib().ihpr_softargmax3d_bwd(heat.data_ptr(), _dtype_code(heat), B, J, D, H, W, coords.data_ptr(),
                                              stats.data_ptr(), g.data_ptr(), grad_heat.data_ptr(), stream))
        return grad_heat, None


class _IntegralL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heat, gt, vis, hd):
        heat = heat.contiguous()
        J = gt.shape[1]
        coords, stats, loss = _fwd(heat, J, (gt, vis, hd))
        ctx.joint_num = J
        ctx.variant = lib().ihpr_get_variant()
        ctx.save_for_backward(heat, coords, stats, gt, vis, hd)
        ctx.mark_non_differentiable(coords)
        ctx.set_materialize_grads(False)         # no zero-fill launch for the unused gradient of `coords`
        return loss, coords

    @staticmethod
    def backward(ctx, grad_loss, _grad_coords):
        if grad_loss is None:
            return None, None, None, None
        heat, coords, stats, gt, vis, hd = ctx.saved_tensors
        J = ctx.joint_num
        B, D, H, W = _shape(heat, J)
        go = grad_loss.to(torch.float32).contiguous()
        grad_heat = torch.empty_like(heat)
        with _on_device(heat.device) as stream:
            lib().ihpr_set_variant(ctx.variant)
            check(lib().ihpr_integral_l1_bwd(heat.data_ptr(), _dtype_code(heat), B, J, D, H, W, coords.data_ptr(),
                                             stats.data_ptr(), gt.data_ptr(), vis.data_ptr(), hd.data_ptr(),
                                             go.data_ptr(), grad_heat.data_ptr(), stream))
        return grad_heat, None, None, None


def _fused_fwd_bwd(heat, gt, vis, hd):
    """K5 through the C-ABI: (loss, coords, stats, d loss / d heat for upstream gradient 1) in one launch."""
    J = gt.shape[1]
    B, D, H, W = _shape(heat, J)
    dev = heat.device
    R = B * J
    with _on_device(dev) as stream:
        _, ws_ptr, ws_n = _planned_workspace(dev, stream, B, J, D, H, W)
        small = torch.empty(R * 5 + 1, dtype=torch.float32, device=dev)
        base = small.data_ptr()
        grad_unit = torch.empty_like(heat)
        check(lib().ihpr_integral_l1_fwd_bwd(heat.data_ptr(), _dtype_code(heat), B, J, D, H, W, gt.data_ptr(), vis.data_ptr(),
                                             hd.data_ptr(), base + R * 20, base, base + R * 12,
                                             grad_unit.data_ptr(), ws_ptr, ws_n, stream))
    return small[R * 5:].view(()), small[:R * 3].view(B, J, 3), small[R * 3:R * 5].view(B, J, 2), grad_unit


class _IntegralL1Fused(torch.autograd.Function):
    """Loss and d loss / d heat in one launch (K5); backward only rescales by the upstream gradient."""

    @staticmethod
    def forward(ctx, heat, gt, vis, hd):
        heat = heat.contiguous()
        J = gt.shape[1]
        loss, coords, stats, grad_unit = _fused_fwd_bwd(heat, gt, vis, hd)
        ctx.variant = lib().ihpr_get_variant()
        ctx.joint_num = J
        ctx.dtype_code = _dtype_code(heat)
        ctx.grad_unit = grad_unit           # consumed (scaled in place) by the first backward
        ctx.save_for_backward(heat, coords, stats, gt, vis, hd)
        ctx.mark_non_differentiable(coords)
        ctx.set_materialize_grads(False)         # no zero-fill launch for the unused gradient of `coords`
        return loss, coords

    @staticmethod
    def backward(ctx, grad_loss, _grad_coords):
        if grad_loss is None:
            return None, None, None, None
        grad_heat = ctx.grad_unit
        go = grad_loss if grad_loss.dtype == torch.float32 else grad_loss.to(torch.float32)
        L = lib()
        if grad_heat is not None:           # the usual case, kept short: it runs on autograd's thread between two launches
            ctx.grad_unit = None
            dev = grad_heat.device
            if torch.cuda.current_device() != dev.index:
                torch.cuda.set_device(dev)
            check(L.ihpr_scale_grad(grad_heat.data_ptr(), ctx.dtype_code, grad_heat.numel(), go.data_ptr(), _raw_stream(dev.index)))
            return grad_heat, None, None, None
        heat, coords, stats, gt, vis, hd = ctx.saved_tensors
        J = ctx.joint_num
        B, D, H, W = _shape(heat, J)
        with _on_device(heat.device) as stream:
            # a second backward through a retained graph: the unit gradient is gone, recompute with K2
            L.ihpr_set_variant(ctx.variant)
            grad_heat = torch.empty_like(heat)
            check(L.ihpr_integral_l1_bwd(heat.data_ptr(), _dtype_code(heat), B, J, D, H, W, coords.data_ptr(),
                                         stats.data_ptr(), gt.data_ptr(), vis.data_ptr(), hd.data_ptr(),
                                         go.data_ptr(), grad_heat.data_ptr(), stream))
        return grad_heat, None, None, None


def _normalise(heat):
    # fp16 / fp64 heatmaps are computed in fp32 (the cast is differentiable torch plumbing)
    if heat.dtype not in (torch.float32, torch.bfloat16):
        if not heat.is_floating_point():
            raise TypeError("heatmaps must be floating point, got %s" % heat.dtype)
        return heat.to(torch.float32)
    return heat


def soft_argmax(heatmaps, joint_num):
    """(B, J*D, H, W) heatmaps -> (B, J, 3) expected (x, y, z) voxel coordinates, fp32.
    Same contract as /root/reference/common/nets/loss.py:13-34; D is inferred as C // joint_num."""
    if isinstance(heatmaps, DeferredHeatmap):
        d = heatmaps
        if torch.is_grad_enabled() and d.requires_grad:           # K3 is forward-only: the differentiable route is conv + K1
            return soft_argmax(d.materialize(), joint_num)
        return fused_head_soft_argmax(d.feat, d.weight, d.bias, joint_num)
    assert isinstance(heatmaps, torch.Tensor)                    # loss.py:14
    _require_cuda(heatmaps, "heatmaps")
    _shape(heatmaps, joint_num)
    if heatmaps.shape[0] == 0:                                   # empty batch: the reference returns an empty (0, J, 3) tensor
        return heatmaps.new_zeros((0, joint_num, 3), dtype=torch.float32) + 0.0 * heatmaps.sum()
    if not (torch.is_grad_enabled() and heatmaps.requires_grad):
        # inference (main/test.py:53-65 runs under no_grad): no autograd node, no stats for a backward that never comes
        return _fwd(_normalise(heatmaps.detach()).contiguous(), int(joint_num), want_stats=False)[0]
    return _SoftArgmax3D.apply(_normalise(heatmaps), int(joint_num))


def integral_l1_loss(heatmap_out, gt_coord, gt_vis, gt_have_depth, return_coords=False, fused_backward=None):
    """Fused soft-argmax + L1 coordinate loss (loss.py:36-52).

    fused_backward=True (default when the heatmaps require grad; IHPR_FUSED=0 disables): the forward launch also
    produces d loss / d heat (K5, DRAM traffic 2V) and backward() only applies the upstream gradient.
    fused_backward=False: one forward launch (K1) and one recomputing backward launch (K2), traffic 3V."""
    if isinstance(heatmap_out, DeferredHeatmap):
        d = heatmap_out
        return fused_head_integral_l1_loss(d.feat, d.weight, d.bias, gt_coord, gt_vis, gt_have_depth, return_coords=return_coords)
    _require_cuda(heatmap_out, "heatmap_out")
    if gt_coord.dim() != 3 or gt_coord.shape[2] != 3:
        raise ValueError("gt_coord must be (B, J, 3), got %s" % (tuple(gt_coord.shape),))
    B, J = gt_coord.shape[0], gt_coord.shape[1]
    _shape(heatmap_out, J)
    if heatmap_out.shape[0] != B:
        raise ValueError("batch mismatch: heatmaps %d vs gt_coord %d" % (heatmap_out.shape[0], B))
    if B == 0:                                                   # loss.py:52: mean() of an empty tensor is NaN
        loss = heatmap_out.sum() * float("nan")
        return (loss, heatmap_out.new_zeros((0, J, 3), dtype=torch.float32)) if return_coords else loss
    dev = heatmap_out.device
    gt = _f32(gt_coord, dev, (B, J, 3), "gt_coord")
    vis = _f32(gt_vis, dev, (B, J), "gt_vis")
    hd = _f32(gt_have_depth, dev, (B, 1), "gt_have_depth")
    if fused_backward is None:
        fused_backward = os.environ.get("IHPR_FUSED", "1") != "0"
    heat = _normalise(heatmap_out)
    if fused_backward and heat.requires_grad and torch.is_grad_enabled():
        loss, coords = _IntegralL1Fused.apply(heat, gt, vis, hd)
    else:
        loss, coords = _IntegralL1.apply(heat, gt, vis, hd)
    return (loss, coords) if return_coords else loss


def integral_l1_step(heatmap_out, gt_coord, gt_vis, gt_have_depth):
    """criterion + ``loss.backward()`` of main/train.py:67-71 as ONE call and one launch (K5): computes the loss and hands
    ``d loss / d heat`` straight to autograd -- ``heatmap_out.backward(grad)`` for the output of a network, ``.grad`` for a leaf --
    so the step has neither the ones-fill nor the rescale launch that ``loss.backward()`` costs.  Returns ``(loss, coords)``,
    both detached.  Same numbers as ``integral_l1_loss(...)`` followed by ``loss.backward()``."""
    if isinstance(heatmap_out, DeferredHeatmap):
        loss, coords = integral_l1_loss(heatmap_out, gt_coord, gt_vis, gt_have_depth, return_coords=True)
        loss.backward()
        return loss.detach(), coords
    _require_cuda(heatmap_out, "heatmap_out")
    if gt_coord.dim() != 3 or gt_coord.shape[2] != 3:
        raise ValueError("gt_coord must be (B, J, 3), got %s" % (tuple(gt_coord.shape),))
    B, J = gt_coord.shape[0], gt_coord.shape[1]
    _shape(heatmap_out, J)
    if heatmap_out.shape[0] != B or B == 0:
        raise ValueError("batch mismatch or empty batch: heatmaps %d vs gt_coord %d" % (heatmap_out.shape[0], B))
    if not heatmap_out.requires_grad:
        raise IhprError("integral_l1_step needs heatmaps that require grad (use integral_l1_loss for a forward-only loss)")
    if heatmap_out.dtype not in (torch.float32, torch.bfloat16) or not heatmap_out.is_contiguous():
        loss, coords = integral_l1_loss(heatmap_out, gt_coord, gt_vis, gt_have_depth, return_coords=True)      # cast / copy: let autograd route it
        loss.backward()
        return loss.detach(), coords
    dev = heatmap_out.device
    gt = _f32(gt_coord, dev, (B, J, 3), "gt_coord")
    vis = _f32(gt_vis, dev, (B, J), "gt_vis")
    hd = _f32(gt_have_depth, dev, (B, 1), "gt_have_depth")
    loss, coords, _, grad = _fused_fwd_bwd(heatmap_out.detach(), gt, vis, hd)
    if heatmap_out.is_leaf:
        if heatmap_out.grad is None:
            heatmap_out.grad = grad
        else:
            heatmap_out.grad += grad
    else:
        heatmap_out.backward(grad)
    return loss, coords


def integral_l1_fwd_bwd_host(heat, gt_coord, gt_vis, gt_have_depth, grad_out=1.0, want_grad=True, device=0, slices=8,
                             out=None):
    """Host-buffer step (ihpr_integral_l1_fwd_bwd_host): CPU tensors in, CPU tensors out, the copies
    are inside.  `heat` / `out` should be pinned for the copies to overlap with the kernels."""
    for t, n in ((heat, "heat"), (gt_coord, "gt_coord"), (gt_vis, "gt_vis"), (gt_have_depth, "gt_have_depth")):
        if t.is_cuda:
            raise ValueError("%s must be a host tensor for the *_host entry point" % n)
    heat = heat.contiguous()
    B, J = gt_coord.shape[0], gt_coord.shape[1]
    _, D, H, W = _shape(heat, J)
    gt = gt_coord.to(torch.float32).contiguous()
    vis = gt_vis.to(torch.float32).contiguous()
    hd = gt_have_depth.to(torch.float32).contiguous()
    if out is None:
        out = {}
    loss = out.get("loss")
    if loss is None:
        loss = out["loss"] = torch.empty(1, dtype=torch.float32)
    coords = out.get("coords")
    if coords is None:
        coords = out["coords"] = torch.empty((B, J, 3), dtype=torch.float32)
    grad = None
    if want_grad:
        grad = out.get("grad")
        if grad is None:
            grad = out["grad"] = torch.empty_like(heat)
    check(lib().ihpr_integral_l1_fwd_bwd_host(heat.data_ptr(), _dtype_code(heat), B, J, D, H, W, gt.data_ptr(),
                                              vis.data_ptr(), hd.data_ptr(), float(grad_out), loss.data_ptr(),
                                              coords.data_ptr(), grad.data_ptr() if grad is not None else None,
                                              int(device), int(slices)))
    return loss, coords, grad


def fused_head_soft_argmax(x, weight, bias, joint_num, return_stats=False):
    """coords of soft_argmax(conv1x1(x, weight, bias)) in ONE launch (K3: tcgen05 GEMM + soft-argmax epilogue); the
    (B, J*D, H, W) heat-map is never written.  Forward only -- the inference tail of main/test.py:62-65.
    x: (B, K, H, W) cuda tensor (made bf16 / channels_last if it is not), weight: (J*D, K[, 1, 1]), bias: (J*D)."""
    _require_cuda(x, "x")
    if x.dim() != 4:
        raise ValueError("x must be (B, K, H, W), got %s" % (tuple(x.shape),))
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        raise IhprError("fused_head_soft_argmax is forward-only (use it under torch.no_grad(); training uses conv + K5)")
    B, K, H, W = x.shape
    M = weight.shape[0]
    if joint_num <= 0 or M % joint_num != 0:
        raise ValueError("%d output channels are not a multiple of joint_num %d" % (M, joint_num))
    D = M // joint_num
    xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)      # physical (B, H, W, K)
    wb = weight.detach().reshape(M, K).to(torch.bfloat16).contiguous()
    bf = (bias.detach() if bias is not None else torch.zeros(M, device=x.device)).to(torch.float32).contiguous()
    dev = x.device
    coords = torch.empty((B, joint_num, 3), dtype=torch.float32, device=dev)
    stats = torch.empty((B, joint_num, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib().ihpr_head_softargmax_fwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr(), B, K, joint_num, D, H, W,
                                             coords.data_ptr(), stats.data_ptr(), stream))
    return (coords, stats) if return_stats else coords


def deconv_bn_relu(x, weight, bn_weight, bn_bias, running_mean, running_var, eps=1e-5):
    """relu(batch_norm(conv_transpose2d(x, weight, stride=2, padding=1))) with the BatchNorm in eval mode, as ONE tensor-core kernel
    (K9, csrc/deconv_bn_relu.cu): the second / third deconv block of HeadNet at inference (main/model.py:22-38).  x: (B, C_in, H, W)
    cuda tensor with W = 32 (H % 8 == 0) or W = 16 (H % 16 == 0), made bf16 / channels_last if it is not; weight: the ConvTranspose2d
    weight (C_in, 256, 4, 4).  Returns the (B, 256, 2 H, 2 W) bf16 channels_last activation -- the operand fused_head_soft_argmax reads without a copy.  Forward only.
    The re-laid weights and the folded BatchNorm are cached per (device, stream, weight tensor) for as long as the SAME parameter tensor
    objects are passed unchanged (identity through weak references + in-place version counters: a module's parameters and buffers), so a
    steady inference loop is one launch per call; temporaries are prepared every time."""
    _require_cuda(x, "x")
    if x.dim() != 4 or weight.dim() != 4 or tuple(weight.shape[2:]) != (4, 4) or weight.shape[0] != x.shape[1]:
        raise ValueError("x must be (B, C_in, H, W) and weight (C_in, C_out, 4, 4), got %s / %s" % (tuple(x.shape), tuple(weight.shape)))
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        raise IhprError("deconv_bn_relu is forward-only (inference with running statistics); call it under torch.no_grad()")
    B, Cin, H, W = x.shape
    Cout = weight.shape[1]
    dev = x.device
    if B == 0:                                      # an empty batch is an empty result, as with the stock modules
        return torch.empty((0, Cout, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    params = (weight, bn_weight, bn_bias, running_mean, running_var)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream, id(weight))      # one prepared set per layer
        stamp = tuple((t.data_ptr(), t._version, t.dtype) for t in params) + (float(eps), Cin, Cout)
        entry = _DECONV_PREPARED.get(key)
        # valid only for the very same tensor OBJECTS, unchanged: a new tensor can land on a recycled address with the same version counter
        if entry is not None and not (entry[0] == stamp and all(r() is t for r, t in zip(entry[2], params))):
            entry = None
        if entry is None:
            if len(_DECONV_PREPARED) >= 32:             # temporaries as parameters (tests, one-off calls) must not pile up workspaces
                _DECONV_PREPARED.clear()
            nbytes = lib().ihpr_deconv_bn_relu_workspace_bytes(Cin, Cout)
            if nbytes == 0:
                raise IhprError("deconv_bn_relu: bad channel counts %d -> %d" % (Cin, Cout))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            wb = weight.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
            g, b_, m, v = (t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in params[1:])
            check(lib().ihpr_deconv_bn_relu_prepare(wb.data_ptr(), g.data_ptr(), b_.data_ptr(), m.data_ptr(), v.data_ptr(), float(eps), Cin, Cout,
                                                    ws.data_ptr(), nbytes, stream))
            entry = (stamp, ws, tuple(weakref.ref(t) for t in params))
            _DECONV_PREPARED[key] = entry
        y = torch.empty((B, Cout, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        check(lib().ihpr_deconv_bn_relu(xb.data_ptr(), entry[1].data_ptr(), B, Cin, Cout, H, W, y.data_ptr(), stream))
    return y


_DECONV_PREPARED = {}


def _nhwc_bf16(t):
    """(B, C, H, W) tensor -> bf16 with channels_last strides, i.e. physically (B, H, W, C) dense; no copy when it already is."""
    return t.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


class _DeconvBnReluTrain(torch.autograd.Function):
    """One deconv block of HeadNet in training (main/model.py:22-38): ConvTranspose2d(256, 256, k4, s2, p1, bias=False) + BatchNorm2d with
    batch statistics + ReLU.  Forward: the K9 GEMM in its training mode (raw output + per-channel sums from the epilogue), a finalize launch,
    one normalise + ReLU pass.  Backward: BatchNorm + ReLU backward in two streaming passes (K10), the input gradient as one tensor-core GEMM
    (K9, mode kDgrad), the weight gradient as 16 more (K11: contraction over the pixels, MN-major operands) -- no library GEMM.
    Saved for backward: the bf16 input, the bf16 raw convolution output and 4 x 256 floats."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, momentum, eps):
        B, Cin, H, W = x.shape
        Cout = weight.shape[1]
        dev = x.device
        xb = _nhwc_bf16(x)
        if weight.dtype == torch.bfloat16 and weight.is_contiguous():
            wb = weight.detach()
        else:           # cast and re-layout (a channels_last parameter) in ONE copy kernel
            wb = torch.empty(weight.shape, dtype=torch.bfloat16, device=dev)
            wb.copy_(weight.detach())
        g = gamma.detach().to(torch.float32).contiguous()
        b_ = beta.detach().to(torch.float32).contiguous()
        for t, n in ((running_mean, "running_mean"), (running_var, "running_var")):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev):
                raise IhprError("%s must be a contiguous fp32 tensor on %s (it is updated in place)" % (n, dev))
        L = lib()
        y_raw = torch.empty((B, Cout, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        out = torch.empty_like(y_raw)
        saved = torch.empty((4, Cout), dtype=torch.float32, device=dev)
        nbytes = L.ihpr_deconv_train_workspace_bytes(Cin, Cout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with _on_device(dev) as stream:
            check(L.ihpr_deconv_bn_relu_train_fwd(xb.data_ptr(), wb.data_ptr(), g.data_ptr(), b_.data_ptr(),
                                                  running_mean.data_ptr() if running_mean is not None else None,
                                                  running_var.data_ptr() if running_var is not None else None, float(momentum), float(eps),
                                                  B, Cin, Cout, H, W, y_raw.data_ptr(), out.data_ptr(), saved.data_ptr(), ws.data_ptr(), nbytes, stream))
        if running_mean is not None:
            # written through raw pointers: tell autograd (and K9's prepared-parameter cache, which stamps the version counters)
            torch.autograd.graph.increment_version(running_mean)
            torch.autograd.graph.increment_version(running_var)
        ctx.save_for_backward(xb, wb, y_raw, saved)
        ctx.meta = (x.dtype, weight.dtype, gamma.dtype, beta.dtype)
        ctx.variant = L.ihpr_get_variant()          # per calling thread: autograd runs backward on another one
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xb, wb, y_raw, saved = ctx.saved_tensors
        x_dtype, w_dtype, g_dtype, b_dtype = ctx.meta
        B, Cin, H, W = xb.shape
        Cout = wb.shape[1]
        dev = xb.device
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        L = lib()
        L.ihpr_set_variant(ctx.variant)
        go = _nhwc_bf16(grad_out)
        dy = torch.empty_like(y_raw)
        dgb = torch.empty((2, Cout), dtype=torch.float32, device=dev)
        dx = torch.empty((B, Cin, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last) if need_x else None
        nbytes = L.ihpr_deconv_train_workspace_bytes(Cin, Cout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with _on_device(dev) as stream:
            check(L.ihpr_deconv_bn_relu_train_bwd(go.data_ptr(), y_raw.data_ptr(), wb.data_ptr(), saved.data_ptr(), B, Cin, Cout, H, W, dy.data_ptr(),
                                                  dgb[0].data_ptr(), dgb[1].data_ptr(), dx.data_ptr() if dx is not None else None,
                                                  ws.data_ptr(), nbytes, stream))
        dw = None
        if need_w and os.environ.get("IHPR_DECONV_WGRAD", "k11") == "library":
            # comparison arm: the weight gradient from the library (cuDNN wgrad through aten)
            dw = torch.ops.aten.convolution_backward(dy, xb, wb, None, [2, 2], [1, 1], [1, 1], True, [0, 0], 1, [False, True, False])[1].to(w_dtype)
        elif need_w:
            # K11: x^T . dy per tap on the tensor cores, fp32 partials added in a fixed order
            dw = torch.empty((Cin, Cout, 4, 4), dtype=torch.float32, device=dev)
            nb2 = L.ihpr_deconv_wgrad_workspace_bytes(Cin, Cout)
            ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
            with _on_device(dev) as stream:
                check(L.ihpr_deconv_wgrad(xb.data_ptr(), dy.data_ptr(), B, Cin, Cout, H, W, dw.data_ptr(), ws2.data_ptr(), nb2, stream))
            dw = dw.to(w_dtype)
        return (dx.to(x_dtype) if dx is not None else None, dw, dgb[0].to(g_dtype) if ctx.needs_input_grad[2] else None,
                dgb[1].to(b_dtype) if ctx.needs_input_grad[3] else None, None, None, None, None)


def deconv_bn_relu_train(x, weight, bn_weight, bn_bias, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    """relu(batch_norm(conv_transpose2d(x, weight, stride=2, padding=1), training=True)) for the 256 -> 256 deconv blocks of HeadNet
    (main/model.py:22-38) with autograd: the forward GEMM (tcgen05) produces the batch statistics in its epilogue, BatchNorm + ReLU and
    their backward are streaming kernels, the input gradient is a second tcgen05 GEMM.  x: (B, 256, H, W) cuda tensor with W = 32
    (H % 8 == 0) or W = 16 (H % 16 == 0); returns the (B, 256, 2 H, 2 W) bf16 channels_last activation.  ``running_mean`` / ``running_var``
    (fp32) are updated in place like torch.nn.BatchNorm2d does; the caller increments ``num_batches_tracked``."""
    _require_cuda(x, "x")
    if x.dim() != 4 or weight.dim() != 4 or tuple(weight.shape[2:]) != (4, 4) or weight.shape[0] != x.shape[1]:
        raise ValueError("x must be (B, C_in, H, W) and weight (C_in, C_out, 4, 4), got %s / %s" % (tuple(x.shape), tuple(weight.shape)))
    if x.shape[0] == 0:
        raise ValueError("batch statistics of an empty batch do not exist")
    if momentum is None:
        raise ValueError("momentum=None (cumulative moving average) is not supported by the fused block")
    return _DeconvBnReluTrain.apply(x, weight, bn_weight, bn_bias, running_mean, running_var, momentum, eps)


class _FusedHeadIntegralL1(torch.autograd.Function):
    """final_layer (1x1 conv) + soft-argmax + L1 loss with the heat-map living only in TMEM: K3 forward; K4w / K4x backward
    (dW, dbias, dX in-kernel -- no heat-map gradient in HBM, no library GEMM).
    Saved for backward: the bf16 activations / weight and 5 floats per joint -- no heat-map, no softmax."""

    @staticmethod
    def forward(ctx, x, weight, bias, gt, vis, hd):
        B, K, H, W = x.shape
        M, J = weight.shape[0], gt.shape[1]
        xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wb = weight.detach().reshape(M, K).to(torch.bfloat16).contiguous()
        bf = bias.detach().to(torch.float32).contiguous()
        dev = x.device
        coords = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
        stats = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            check(lib().ihpr_head_softargmax_fwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr(), B, K, J, M // J, H, W,
                                                 coords.data_ptr(), stats.data_ptr(), stream))
            # loss.py:49-52 on the (B, J, 3) coordinates: one small launch
            loss = torch.empty((), dtype=torch.float32, device=dev)
            check(lib().ihpr_integral_l1_from_coords(coords.data_ptr(), gt.data_ptr(), vis.data_ptr(), hd.data_ptr(), B, J, loss.data_ptr(), stream))
        ctx.save_for_backward(xb, wb, bf, coords, stats, gt, vis, hd)
        ctx.variant = lib().ihpr_get_variant()
        ctx.meta = (x.dtype, weight.dtype, bias.dtype, tuple(weight.shape), tuple(weight.stride()))
        ctx.mark_non_differentiable(coords)
        ctx.set_materialize_grads(False)
        return loss, coords

    @staticmethod
    def backward(ctx, grad_loss, _grad_coords):
        if grad_loss is None:
            return None, None, None, None, None, None
        xb, wb, bf, coords, stats, gt, vis, hd = ctx.saved_tensors
        x_dtype, w_dtype, b_dtype, w_shape, w_stride = ctx.meta
        B, K, H, W = xb.shape
        M, J = wb.shape[0], gt.shape[1]
        N = H * W
        go = grad_loss.to(torch.float32).contiguous()
        dev = xb.device
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        L = lib()
        if ctx.variant == 6:
            # comparison arm only (ihpr_set_variant(6)): K4 writes the bf16 heat-map gradient and library GEMMs turn it into dX / dW
            dheat = torch.empty((B, M, N), dtype=torch.bfloat16, device=dev)        # d loss / d heat-map, (B, J*D, H, W) layout
            db_part = torch.empty((B, 4, M), dtype=torch.float32, device=dev)
            with _on_device(dev) as stream:
                check(L.ihpr_head_integral_l1_bwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr(), B, K, J, M // J, H, W,
                                                  coords.data_ptr(), stats.data_ptr(), gt.data_ptr(), vis.data_ptr(),
                                                  hd.data_ptr(), go.data_ptr(), dheat.data_ptr(), db_part.data_ptr(), stream))
            xn = xb.permute(0, 2, 3, 1).reshape(B, N, K)                              # view of the channels_last activations
            dx = torch.matmul(dheat.transpose(1, 2), wb).view(B, H, W, K).permute(0, 3, 1, 2).to(x_dtype) if need_x else None
            dw = torch.bmm(dheat, xn, out_dtype=torch.float32).sum(0).view(w_shape).to(w_dtype) if need_w else None
            if dw is not None and len(w_shape) == 4 and w_shape[2:] == (1, 1):
                dw = torch.as_strided(dw, w_shape, w_stride)
            db = db_part.sum(dim=(0, 1)).to(b_dtype) if need_b else None
            return dx, dw, db, None, None, None
        # K4w / K4x: dW, dbias and dX straight out of the tensor-core kernels; neither the heat-map nor its gradient is ever stored
        dx = torch.empty((B, H, W, K), dtype=torch.bfloat16, device=dev) if need_x else None       # = (B, K, H, W) channels_last
        dw = torch.empty((M, K), dtype=torch.float32, device=dev) if need_w else None
        db = torch.empty((M,), dtype=torch.float32, device=dev) if need_b else None
        if not (need_x or need_w or need_b):
            return None, None, None, None, None, None
        nbytes = L.ihpr_head_bwd_workspace_bytes(B, K, J, M // J, H, W)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ptr = lambda t: t.data_ptr() if t is not None else None     # noqa: E731
        with _on_device(dev) as stream:
            check(L.ihpr_head_integral_l1_bwd_params(xb.data_ptr(), wb.data_ptr(), bf.data_ptr(), B, K, J, M // J, H, W,
                                                     coords.data_ptr(), stats.data_ptr(), gt.data_ptr(), vis.data_ptr(), hd.data_ptr(),
                                                     go.data_ptr(), ptr(dx), ptr(dw), ptr(db), ws.data_ptr(), nbytes, stream))
        if dx is not None:
            dx = dx.permute(0, 3, 1, 2).to(x_dtype)                   # logical (B, K, H, W), channels_last strides, no copy for bf16 x
        if dw is not None:
            # the gradient in the parameter's own strides (a channels_last (M, K, 1, 1) weight has strides (K, 1, K, K)): same memory,
            # and DDP's bucket views then take it without a re-layout copy
            dw = dw.to(w_dtype)
            dw = torch.as_strided(dw, w_shape, w_stride) if len(w_shape) == 4 and w_shape[2:] == (1, 1) else dw.view(w_shape)
        if db is not None:
            db = db.to(b_dtype)
        return dx, dw, db, None, None, None


def fused_head_integral_l1_loss(x, weight, bias, gt_coord, gt_vis, gt_have_depth, return_coords=False):
    """JointLocationLoss(final_layer(x), ...) for training without ever storing the (B, J*D, H, W) heat-map:
    main/model.py:42 + main/train.py:67-71 as tensor-core launches only (K3 forward; K4w + K4x backward: dW / dbias / dX in-kernel)."""
    _require_cuda(x, "x")
    if x.dim() != 4:
        raise ValueError("x must be (B, K, H, W), got %s" % (tuple(x.shape),))
    if gt_coord.dim() != 3 or gt_coord.shape[2] != 3:
        raise ValueError("gt_coord must be (B, J, 3), got %s" % (tuple(gt_coord.shape),))
    B, J = gt_coord.shape[0], gt_coord.shape[1]
    if x.shape[0] != B:
        raise ValueError("batch mismatch: features %d vs gt_coord %d" % (x.shape[0], B))
    if J <= 0 or weight.shape[0] % J != 0:
        raise ValueError("%d output channels are not a multiple of the %d joints in gt_coord" % (weight.shape[0], J))
    dev = x.device
    gt = _f32(gt_coord, dev, (B, J, 3), "gt_coord")
    vis = _f32(gt_vis, dev, (B, J), "gt_vis")
    hd = _f32(gt_have_depth, dev, (B, 1), "gt_have_depth")
    if bias is None:
        bias = torch.zeros(weight.shape[0], device=dev)
    loss, coords = _FusedHeadIntegralL1.apply(x, weight, bias, gt, vis, hd)
    return (loss, coords) if return_coords else loss


class DeferredHeatmap:
    """The output of ``HeadNet.final_layer`` (main/model.py:42) that has NOT been computed: the deconv features plus the 1x1
    conv's parameters.  ``ResPoseNet(fused_head=True, deferred=True)(img)`` returns one, and the drop-in ``soft_argmax`` /
    ``JointLocationLoss`` consume it with the fused tensor-core kernels (K3 / K4), so the reference's own call sequences

        heatmap_out = model(input_img); loss = JointLocationLoss(heatmap_out, joint_img, joint_vis, have_depth)   # train.py:64-67
        heatmap_out = model(input_img); coord_out = soft_argmax(heatmap_out, joint_num)                           # test.py:62-65

    run unchanged while the (B, J*D, H, W) volume never exists.  It answers the shape questions a caller may ask
    (``shape``, ``size()``, ``dim()``, ``device``, ``dtype``); anything else needs ``materialize()`` (a plain conv)."""

    def __init__(self, feat, weight, bias, joint_num):
        if feat.dim() != 4:
            raise ValueError("features must be (B, K, H, W), got %s" % (tuple(feat.shape),))
        if weight.shape[1] != feat.shape[1]:
            raise ValueError("weight has %d input channels, features have %d" % (weight.shape[1], feat.shape[1]))
        self.feat, self.weight, self.bias, self.joint_num = feat, weight, bias, int(joint_num)

    @property
    def shape(self):
        return torch.Size((self.feat.shape[0], self.weight.shape[0], self.feat.shape[2], self.feat.shape[3]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    device = property(lambda self: self.feat.device)
    dtype = property(lambda self: self.feat.dtype)
    is_cuda = property(lambda self: self.feat.is_cuda)
    requires_grad = property(lambda self: self.feat.requires_grad or self.weight.requires_grad)

    def materialize(self):
        """The heat-map itself, for callers that really need the tensor (differentiable, stock conv)."""
        w = self.weight.reshape(self.weight.shape[0], self.weight.shape[1], 1, 1)
        return torch.nn.functional.conv2d(self.feat, w.to(self.feat.dtype), None if self.bias is None else self.bias.to(self.feat.dtype))


def flip_perm(joint_num, flip_pairs):
    """The pairwise left/right swaps of main/test.py:74-75, applied in order, as one permutation:
    after the swaps joint j holds what joint perm[j] held before."""
    perm = list(range(joint_num))
    for a, b in flip_pairs:
        perm[a], perm[b] = perm[b], perm[a]
    return perm


_perm_cache = {}


def _perm_tensor(joint_num, flip_pairs, dev):
    key = (dev, joint_num, tuple((int(a), int(b)) for a, b in flip_pairs))
    t = _perm_cache.get(key)
    if t is None:
        t = _perm_cache[key] = torch.tensor(flip_perm(joint_num, key[2]), dtype=torch.int32, device=dev)
    return t


def _coords_arg(t, name, dev=None):
    _require_cuda(t, name)
    if dev is not None and t.device != dev:
        raise IhprError("%s is on %s but coord_out is on %s" % (name, t.device, dev))
    return t.detach().to(torch.float32).contiguous()


def coords_to_camera(coord_out, bbox=None, center_cam=None, f=None, c=None, root_idx=None, flipped_coord_out=None, flip_pairs=(),
                     depth_dim=64, output_shape=(64, 64), bbox_3d_depth=2000.0, outputs=("merged", "pixel", "cam")):
    """Test-time post-processing of the soft-argmax result in ONE launch (ihpr_coords_to_camera), nothing leaves the device:

    * ``merged``: flip-test merge of ``coord_out`` with ``flipped_coord_out`` (main/test.py:67-76); a copy of ``coord_out`` without one;
    * ``pixel``:  ``warp_coord_to_original`` (common/utils/pose_utils.py:68-75) with per-sample ``bbox`` (B,4) / ``center_cam`` (B,3);
    * ``cam``:    ``pixel2cam`` (pose_utils.py:14-20) with per-sample ``f``, ``c`` (B,2), minus the root joint when ``root_idx`` is
      given (data/Human36M/Human36M.py:226-228).

    ``depth_dim`` / ``output_shape`` (H, W) / ``bbox_3d_depth`` are cfg.depth_dim, cfg.output_shape, cfg.bbox_3d_shape[0]
    (main/config.py:27-29).  Returns a dict of (B, J, 3) fp32 device tensors for the requested ``outputs``."""
    coords = _coords_arg(coord_out, "coord_out")
    if coords.dim() != 3 or coords.shape[2] != 3:
        raise ValueError("coord_out must be (B, J, 3), got %s" % (tuple(coords.shape),))
    B, J, _ = coords.shape
    dev = coords.device
    bad = set(outputs) - {"merged", "pixel", "cam"}
    if bad or not outputs:
        raise ValueError("outputs must be a non-empty subset of merged / pixel / cam, got %r" % (tuple(outputs),))

    def per_sample(t, width, name, needed):
        if t is None:
            if needed:
                raise ValueError("%s is required for the requested outputs" % name)
            return None
        t = _coords_arg(torch.as_tensor(t, device=dev) if not isinstance(t, torch.Tensor) else t, name, dev)
        if t.dim() == 1:                      # one camera / box for the whole batch
            t = t.expand(B, width).contiguous()
        if tuple(t.shape) != (B, width):
            raise ValueError("%s has shape %s, expected (%d, %d)" % (name, tuple(t.shape), B, width))
        return t

    geo = "pixel" in outputs or "cam" in outputs
    bbox = per_sample(bbox, 4, "bbox", geo)
    center_cam = per_sample(center_cam, 3, "center_cam", geo)
    f = per_sample(f, 2, "f", "cam" in outputs)
    c = per_sample(c, 2, "c", "cam" in outputs)
    flipped = perm = None
    if flipped_coord_out is not None:
        flipped = _coords_arg(flipped_coord_out, "flipped_coord_out", dev)
        if flipped.shape != coords.shape:
            raise ValueError("flipped_coord_out has shape %s, expected %s" % (tuple(flipped.shape), tuple(coords.shape)))
        if len(flip_pairs):
            perm = _perm_tensor(J, flip_pairs, dev)
    root = -1 if root_idx is None else int(root_idx)
    if root >= J:
        raise ValueError("root_idx %d out of range for %d joints" % (root, J))
    out = {k: torch.empty(B, J, 3, dtype=torch.float32, device=dev) for k in outputs}
    if B == 0:
        return out
    ptr = lambda t: t.data_ptr() if t is not None else None     # noqa: E731
    with torch.cuda.device(dev):
        check(lib().ihpr_coords_to_camera(ptr(coords), ptr(flipped), ptr(perm), B, J, int(depth_dim), int(output_shape[0]), int(output_shape[1]),
                                          ptr(bbox), ptr(center_cam), ptr(f), ptr(c), float(bbox_3d_depth), root,
                                          ptr(out.get("merged")), ptr(out.get("pixel")), ptr(out.get("cam")),
                                          torch.cuda.current_stream(dev).cuda_stream))
    return out


def flip_merge(coord_out, flipped_coord_out, width, flip_pairs):
    """Flip-test merge of main/test.py:67-76 on the (B, J, 3) device coordinates, one launch, no host round trip:
    mirror x of the flipped pass (x' = W - x - 1), swap the left/right joints, average with the un-flipped pass."""
    return coords_to_camera(coord_out, flipped_coord_out=flipped_coord_out, flip_pairs=flip_pairs, output_shape=(1, int(width)),
                            outputs=("merged",))["merged"]


def last_launch_count():
    return lib().ihpr_last_launch_count()


def last_path_choice():
    """What the last ``ihpr_integral_l1_fwd_bwd`` of this thread ran: 1 = one launch (K5 / K5c), 2 = K1 + K2; +16 when that call
    was the one that measured both forms for this (device, shape)."""
    return lib().ihpr_last_path_choice()


def set_variant(v):
    check(lib().ihpr_set_variant(int(v)))


def get_variant():
    return lib().ihpr_get_variant()
