"""oracle/augment_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (numpy, no cv2) of what the reference's DatasetLoader.__getitem__ does to one training sample
(/root/reference/data/dataset.py:84-152 and its helpers :184-262, transform from common/base.py:93-95):

  get_aug_config            dataset.py:184-199   scale / rotation / flip / colour factors from the global numpy + random RNGs
  gen_trans                 dataset.py:229-257   3-point affine patch transform (the reference calls cv2.getAffineTransform)
  warp_affine_linear_u8     dataset.py:215       cv2.warpAffine(..., INTER_LINEAR) on a uint8 image, restated from OpenCV's published
                                                 fixed-point algorithm (imgproc imgwarp: inverse matrix in double, 10-bit coordinates,
                                                 5-bit sub-pixel fractions, 15-bit bilinear weights, constant zero border)
  get_item                  dataset.py:84-152    patch + colour scale + clip + ToTensor/Normalize; joints flipped, warped, depth
                                                 normalised, visibility-tested, scaled to heat-map space

OpenCV is a third-party dependency of the reference (un-pinned: README names no version; 4.13.0 is installed in the build
container).  oracle/make_golden.py --aug pins this file by running the reference's own DatasetLoader.__getitem__ (which calls the
real cv2) on seeded synthetic images: patches must be bit-identical, joints within 1e-6.
"""
import math
import random

import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE = 1 << AB_BITS
ROUND_DELTA = AB_SCALE // (1 << INTER_BITS) // 2


def get_aug_config(scale_factor=0.25, rot_factor=30, color_factor=0.2):
    scale = np.clip(np.random.randn(), -1.0, 1.0) * scale_factor + 1.0
    rot = np.clip(np.random.randn(), -2.0, 2.0) * rot_factor if random.random() <= 0.6 else 0
    do_flip = random.random() <= 0.5
    lo, hi = 1.0 - color_factor, 1.0 + color_factor
    color_scale = [random.uniform(lo, hi), random.uniform(lo, hi), random.uniform(lo, hi)]
    return scale, rot, do_flip, color_scale


def _rot(v, rad):
    s, c = np.sin(rad), np.cos(rad)
    return np.array([v[0] * c - v[1] * s, v[0] * s + v[1] * c], dtype=np.float32)


def affine_from_3_points(src, dst):
    """2x3 double matrix M with M @ [x, y, 1] = dst for the three src points (what cv2.getAffineTransform solves)."""
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[i, 0:2], a[i, 2] = src[i], 1.0
        a[i + 3, 3:5], a[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def gen_trans(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv=False):
    src_w, src_h = src_width * scale, src_height * scale
    center = np.array([c_x, c_y], dtype=np.float32)
    rad = np.pi * rot / 180
    down = _rot(np.array([0, src_h * 0.5], dtype=np.float32), rad)
    right = _rot(np.array([src_w * 0.5, 0], dtype=np.float32), rad)
    dcenter = np.array([dst_width * 0.5, dst_height * 0.5], dtype=np.float32)
    src = np.stack([center, center + down, center + right]).astype(np.float32)
    dst = np.stack([dcenter, dcenter + np.array([0, dst_height * 0.5], np.float32), dcenter + np.array([dst_width * 0.5, 0], np.float32)])
    dst = dst.astype(np.float32)
    return affine_from_3_points(dst, src) if inv else affine_from_3_points(src, dst)


def invert_affine(m):
    """OpenCV's in-place inversion of the 2x3 forward map inside warpAffine (double precision)."""
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    i = np.zeros((2, 3))
    i[0, 0], i[0, 1], i[1, 0], i[1, 1] = a11, m[0, 1] * (-d), m[1, 0] * (-d), a22
    i[0, 2] = -i[0, 0] * m[0, 2] - i[0, 1] * m[1, 2]
    i[1, 2] = -i[1, 0] * m[0, 2] - i[1, 1] * m[1, 2]
    return i


def warp_affine_linear_u8(img, m, dsize):
    """cv2.warpAffine(img, m, dsize, flags=INTER_LINEAR) for uint8 HxWxC, borderMode constant 0."""
    w_out, h_out = dsize
    h, w = img.shape[:2]
    im = invert_affine(np.asarray(m, np.float64))
    xs = np.arange(w_out)
    adelta = np.rint(im[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(im[1, 0] * xs * AB_SCALE).astype(np.int64)
    ys = np.arange(h_out)
    x0 = np.rint((im[0, 1] * ys + im[0, 2]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    y0 = np.rint((im[1, 1] * ys + im[1, 2]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    X = (x0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx, sy = X >> INTER_BITS, Y >> INTER_BITS
    ax, ay = X & 31, Y & 31
    src = img.astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return v * ok[..., None]

    w00, w01 = ((32 - ay) * (32 - ax) * 32)[..., None], ((32 - ay) * ax * 32)[..., None]
    w10, w11 = (ay * (32 - ax) * 32)[..., None], (ay * ax * 32)[..., None]
    acc = tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def generate_patch(cvimg, bbox, do_flip, scale, rot, input_shape):
    img = cvimg
    width = img.shape[1]
    c_x, c_y = float(bbox[0] + 0.5 * bbox[2]), float(bbox[1] + 0.5 * bbox[3])
    if do_flip:
        img = img[:, ::-1, :]
        c_x = width - c_x - 1
    trans = gen_trans(c_x, c_y, float(bbox[2]), float(bbox[3]), input_shape[1], input_shape[0], scale, rot)
    patch = warp_affine_linear_u8(img, trans, (int(input_shape[1]), int(input_shape[0])))
    return patch[:, :, ::-1].astype(np.float32), trans


def get_item(cvimg, bbox, joint_img, joint_vis, flip_pairs, aug, input_shape, output_shape, depth_dim, bbox_3d_depth,
             pixel_mean, pixel_std):
    """-> (img (3, H, W) float32 normalised, joint_img (J, 3) float32 in heat-map space, joint_vis (J, 1) float32, trans)."""
    scale, rot, do_flip, color_scale = aug
    joint_img, joint_vis = np.array(joint_img, np.float64), np.array(joint_vis, np.float64)
    width = cvimg.shape[1]
    patch, trans = generate_patch(cvimg, bbox, do_flip, scale, rot, input_shape)
    for i in range(3):
        patch[:, :, i] = np.clip(patch[:, :, i] * color_scale[i], 0, 255)
    if do_flip:
        joint_img[:, 0] = width - joint_img[:, 0] - 1
        for a, b in flip_pairs:
            joint_img[[a, b]] = joint_img[[b, a]]
            joint_vis[[a, b]] = joint_vis[[b, a]]
    for i in range(len(joint_img)):
        joint_img[i, 0:2] = trans @ np.array([joint_img[i, 0], joint_img[i, 1], 1.0])
        joint_img[i, 2] /= (bbox_3d_depth / 2. * scale)
        joint_img[i, 2] = (joint_img[i, 2] + 1.0) / 2.
        joint_vis[i] *= ((joint_img[i, 0] >= 0) & (joint_img[i, 0] < input_shape[1]) & (joint_img[i, 1] >= 0) &
                         (joint_img[i, 1] < input_shape[0]) & (joint_img[i, 2] >= 0) & (joint_img[i, 2] < 1))
    joint_img[:, 0] = joint_img[:, 0] / input_shape[1] * output_shape[1]
    joint_img[:, 1] = joint_img[:, 1] / input_shape[0] * output_shape[0]
    joint_img[:, 2] = joint_img[:, 2] * depth_dim
    chw = np.ascontiguousarray(patch.transpose(2, 0, 1))
    mean = np.array(pixel_mean, np.float32).reshape(3, 1, 1)
    std = np.array(pixel_std, np.float32).reshape(3, 1, 1)
    img = (chw - mean) / std
    return img.astype(np.float32), joint_img.astype(np.float32), (joint_vis > 0).astype(np.float32), trans


def synthetic_image(h, w, seed):
    """Deterministic BGR uint8 test image: smooth colour gradients + blocks + noise (exercises edges and flat areas)."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(xx / 17.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], -1)
    for _ in range(12):
        y0, x0 = int(g.integers(0, h - 8)), int(g.integers(0, w - 8))
        img[y0:y0 + int(g.integers(8, h // 3)), x0:x0 + int(g.integers(8, w // 3))] = g.integers(0, 256, 3)
    img += g.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_annotation(h, w, J, seed):
    """bbox (x, y, w, h) and joints (J, 3): pixel x, y inside / slightly outside the box, root-relative depth in mm."""
    g = np.random.default_rng(seed + 1000)
    bw, bh = w * (0.35 + 0.3 * g.random()), h * (0.35 + 0.3 * g.random())
    bx, by = (w - bw) * g.random(), (h - bh) * g.random()
    bbox = np.array([bx, by, bw, bh], np.float32)
    joints = np.stack([bx + bw * (g.random(J) * 1.3 - 0.15), by + bh * (g.random(J) * 1.3 - 0.15), g.normal(0, 450, J)], 1)
    vis = (g.random((J, 1)) > 0.15).astype(np.float64)
    return bbox, joints, vis
