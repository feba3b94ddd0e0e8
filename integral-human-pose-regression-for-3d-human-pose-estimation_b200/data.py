"""Training-sample preparation on the device: the per-sample work of the reference's ``DatasetLoader.__getitem__``
(/root/reference/data/dataset.py:84-152) for a whole batch in two launches (K7 patches, K8 joints).

Host side (this file): draw the augmentation parameters exactly as the reference does (same global RNGs, same order), build
the 2x3 patch transform, pack every per-sample parameter into ONE pinned record and copy it once.  Device side
(csrc/augment.cu): the warp, colour scale, normalisation and the joint transforms.  Image decoding stays on the host
(cv2.imread in the reference, dataset.py:79); what is uploaded is the raw uint8 BGR image.

There is no CPU path: ``augment_batch`` raises on a non-CUDA device.
"""
import ctypes
import random

import numpy as np
import torch

from ._lib import IhprError, check, lib
from .functional import flip_perm

PIXEL_MEAN = (0.485, 0.456, 0.406)      # main/config.py:31-32
PIXEL_STD = (0.229, 0.224, 0.225)


def get_aug_config(scale_factor=0.25, rot_factor=30, color_factor=0.2):
    """dataset.py:184-199: (scale, rot, do_flip, color_scale) drawn from numpy's and random's GLOBAL generators in the
    reference's order, so the same seeds give the same augmentation as the reference's loader."""
    scale = float(np.clip(np.random.randn(), -1.0, 1.0)) * scale_factor + 1.0
    rot = float(np.clip(np.random.randn(), -2.0, 2.0)) * rot_factor if random.random() <= 0.6 else 0
    do_flip = random.random() <= 0.5
    lo, hi = 1.0 - color_factor, 1.0 + color_factor
    return scale, rot, do_flip, [random.uniform(lo, hi) for _ in range(3)]


NO_AUG = (1.0, 0, False, [1.0, 1.0, 1.0])     # dataset.py:88 (test-time / do_augment False)


def gen_trans_batch(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv=False):
    """dataset.py:229-257 for n boxes at once: (n, 2, 3) float64.  The reference builds three float32 point pairs per box (centre,
    centre + down, centre + right; the edge vectors rotated in double, rounded to float32) and lets cv2.getAffineTransform solve
    for the map; with those pairs the solution is  L = D S^-1,  t = dc - L sc  where the columns of S / D are the (down, right)
    edge vectors.  The float32 roundings are kept where the reference has them."""
    f32, f64 = np.float32, np.float64
    c_x, c_y, src_width, src_height, scale, rot = (np.atleast_1d(np.asarray(v, f64)) for v in (c_x, c_y, src_width, src_height, scale, rot))
    rad = np.pi * rot / 180
    sn, cs = np.sin(rad), np.cos(rad)
    half_h = (src_height * scale * 0.5).astype(f32).astype(f64)
    half_w = (src_width * scale * 0.5).astype(f32).astype(f64)
    down = np.stack([-(half_h * sn), half_h * cs], 1).astype(f32)                 # rotate_2d((0, h/2))
    right = np.stack([half_w * cs, half_w * sn], 1).astype(f32)                  # rotate_2d((w/2, 0))
    sc = np.stack([c_x, c_y], 1).astype(f32)
    n = sc.shape[0]
    dc = np.broadcast_to(np.array([dst_width * 0.5, dst_height * 0.5], f32), (n, 2))
    src = np.stack([sc, sc + down, sc + right], 1).astype(f64)                  # (n, 3 points, 2), sums in float32
    dst = np.stack([dc, dc + np.array([0, dst_height * 0.5], f32), dc + np.array([dst_width * 0.5, 0], f32)], 1).astype(f64)
    if inv:
        src, dst = dst, src
    S = np.stack([src[:, 1] - src[:, 0], src[:, 2] - src[:, 0]], 2)            # (n, 2, 2), columns = edge vectors
    D = np.stack([dst[:, 1] - dst[:, 0], dst[:, 2] - dst[:, 0]], 2)
    det = S[:, 0, 0] * S[:, 1, 1] - S[:, 0, 1] * S[:, 1, 0]
    Sinv = np.stack([np.stack([S[:, 1, 1], -S[:, 0, 1]], 1), np.stack([-S[:, 1, 0], S[:, 0, 0]], 1)], 1) / det[:, None, None]
    L = D @ Sinv
    t = dst[:, 0] - np.einsum("nij,nj->ni", L, src[:, 0])
    return np.concatenate([L, t[:, :, None]], 2)


def gen_trans_from_patch(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv=False):
    """dataset.py:229-257 (same argument list): the affine map taking the (scaled, rotated) box around (c_x, c_y) onto the
    dst_width x dst_height patch, as a 2x3 float64 matrix."""
    return gen_trans_batch(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv)[0]


def patch_params_batch(bboxes, img_widths, augs, input_shape):
    """Forward patch transforms (n, 2, 3) for the (mirrored, when flipped) box centres of dataset.py:204-213.  The centre is computed
    in the boxes' own dtype, as the reference's  float(bbox[0] + 0.5 * bbox[2])  does, then carried in double."""
    bboxes = np.asarray(bboxes)
    c_x = (bboxes[:, 0] + 0.5 * bboxes[:, 2]).astype(np.float64)
    c_y = (bboxes[:, 1] + 0.5 * bboxes[:, 3]).astype(np.float64)
    flip = np.array([bool(a[2]) for a in augs])
    c_x = np.where(flip, np.asarray(img_widths, np.float64) - c_x - 1, c_x)
    return gen_trans_batch(c_x, c_y, bboxes[:, 2], bboxes[:, 3], input_shape[1], input_shape[0], [a[0] for a in augs], [a[1] for a in augs])


def patch_params(bbox, img_width, aug, input_shape):
    """One sample of ``patch_params_batch``."""
    return patch_params_batch(np.asarray(bbox)[None], [img_width], [aug], input_shape)[0]


def _record_layout(B, J):
    """Offsets (in bytes) of the fields of the single per-batch parameter record; doubles first (8-byte aligned)."""
    fields = (("trans", np.float64, B * 6), ("scale", np.float64, B), ("joint_img", np.float64, B * J * 3), ("joint_vis", np.float64, B * J),
              ("sizes", np.int32, B * 2), ("do_flip", np.int32, B), ("perm", np.int32, J), ("color", np.float32, B * 3))
    off, layout = 0, {}
    for name, dt, n in fields:
        layout[name] = (off, dt, n)
        off += n * np.dtype(dt).itemsize
    return layout, off


def augment_batch(images, sizes, bboxes, joint_img, joint_vis, augs, flip_pairs=(), input_shape=(256, 256), output_shape=(64, 64),
                  depth_dim=64, bbox_3d_depth=2000.0, pixel_mean=PIXEL_MEAN, pixel_std=PIXEL_STD, channels_last=False):
    """Steps 3-4 of ``DatasetLoader.__getitem__`` + the ToTensor/Normalize transform for a batch.

    images     (B, Hs, Ws, 3) uint8 CUDA tensor: BGR images as cv2.imread returns them, zero-padded to a common size
    sizes      (B, 2) valid (rows, cols) of each image                  [host]
    bboxes     (B, 4) x, y, w, h                                        [host]
    joint_img  (B, J, 3) x, y in image pixels, root-relative depth mm   [host]
    joint_vis  (B, J) or (B, J, 1)                                      [host]
    augs       B tuples (scale, rot, do_flip, color_scale) from ``get_aug_config()`` (or ``NO_AUG``)

    Returns (img_patch (B, 3, H, W) fp32 normalised, joint_img (B, J, 3) fp32 in heat-map space, joint_vis (B, J, 1) fp32) on the
    images' device -- what dataset.py:146-152 returns per sample, batched."""
    if not isinstance(images, torch.Tensor) or not images.is_cuda:
        raise IhprError("ihpr_b200: augment_batch needs the uint8 images on a CUDA device (there is no CPU fallback)")
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
        raise ValueError("images must be (B, Hs, Ws, 3) uint8, got %s %s" % (tuple(images.shape), images.dtype))
    images = images.contiguous()
    dev = images.device
    B, Hs, Ws, _ = images.shape
    sizes = np.asarray(sizes, np.int32).reshape(B, 2)
    if B and (sizes.min() <= 0 or sizes[:, 0].max() > Hs or sizes[:, 1].max() > Ws):
        raise ValueError("sizes must lie inside the padded (%d, %d) image" % (Hs, Ws))
    bboxes = np.asarray(bboxes).reshape(B, 4)          # dtype kept: dataset.py:204-205 does the centre arithmetic in the box's own dtype
    joint_img = np.asarray(joint_img, np.float64)
    if joint_img.ndim != 3 or joint_img.shape[0] != B or joint_img.shape[2] != 3:
        raise ValueError("joint_img must be (B, J, 3), got %s" % (joint_img.shape,))
    J = joint_img.shape[1]
    joint_vis = np.asarray(joint_vis, np.float64).reshape(B, J)
    if len(augs) != B:
        raise ValueError("need one augmentation tuple per sample")
    in_h, in_w = int(input_shape[0]), int(input_shape[1])

    layout, nbytes = _record_layout(B, J)
    host = torch.empty(max(nbytes, 8), dtype=torch.uint8, pin_memory=True)
    raw = host.numpy()

    def field(name):
        off, dt, n = layout[name]
        return raw[off:off + n * np.dtype(dt).itemsize].view(dt)

    if B:
        field("trans")[:] = patch_params_batch(bboxes, sizes[:, 1], augs, (in_h, in_w)).reshape(-1)
    field("scale")[:] = [a[0] for a in augs]
    field("joint_img")[:] = joint_img.reshape(-1)
    field("joint_vis")[:] = joint_vis.reshape(-1)
    field("sizes")[:] = sizes.reshape(-1)
    field("do_flip")[:] = [1 if a[2] else 0 for a in augs]
    field("perm")[:] = flip_perm(J, flip_pairs)
    field("color")[:] = np.asarray([a[3] for a in augs], np.float32).reshape(-1) if B else []
    rec = host.to(dev, non_blocking=True)
    base = rec.data_ptr()
    ptr = {k: base + v[0] for k, v in layout.items()}

    shape = (B, in_h, in_w, 3) if channels_last else (B, 3, in_h, in_w)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    gt_coord = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
    gt_vis = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
    if B:
        mean = (lib_float3(pixel_mean), lib_float3(pixel_std))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            check(lib().ihpr_augment_patches(images.data_ptr(), ptr["sizes"], B, Hs, Ws, ptr["trans"], ptr["do_flip"], ptr["color"], mean[0], mean[1],
                                             in_h, in_w, out.data_ptr(), 1 if channels_last else 0, stream))
            check(lib().ihpr_augment_joints(ptr["joint_img"], ptr["joint_vis"], ptr["sizes"], ptr["trans"], ptr["scale"], ptr["do_flip"], ptr["perm"],
                                            B, J, in_h, in_w, int(output_shape[0]), int(output_shape[1]), int(depth_dim), float(bbox_3d_depth),
                                            gt_coord.data_ptr(), gt_vis.data_ptr(), stream))
    if channels_last:
        out = out.permute(0, 3, 1, 2)          # logical NCHW, channels_last strides
    return out, gt_coord, gt_vis


def lib_float3(values):
    if len(values) != 3:
        raise ValueError("expected 3 per-channel values, got %r" % (values,))
    return (ctypes.c_float * 3)(*[float(v) for v in values])
