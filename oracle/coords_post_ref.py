"""oracle/coords_post_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the reference's test-time post-processing of the soft-argmax result:

* ``flip_merge``            main/test.py:73-76   (inline in the reference's main(): restated with the same torch statements)
* ``warp_to_original``      common/utils/pose_utils.py:68-75
* ``pixel_to_cam``          common/utils/pose_utils.py:14-20
* ``evaluate_sample``       data/Human36M/Human36M.py:203-228: the order and dtypes the reference chains them in
                            (float32 prediction copy, float64 camera array, root-joint alignment)

Pinned by tests/golden/post_*.npz, which oracle/make_golden.py writes by calling the reference's own
``warp_coord_to_original`` / ``pixel2cam`` (tests/test_oracle.py checks this file against them bit for bit).
"""
import numpy as np
import torch


def flip_merge(coord_out, flipped_coord_out, width, flip_pairs):
    f = flipped_coord_out.clone()
    f[:, :, 0] = width - f[:, :, 0] - 1
    for a, b in flip_pairs:
        f[:, a, :], f[:, b, :] = f[:, b, :].clone(), f[:, a, :].clone()
    return (coord_out + f) / 2.


def warp_to_original(joint_out, bbox, center_cam, depth_dim, output_shape, bbox_3d_depth):
    x = joint_out[:, 0] / output_shape[1] * bbox[2] + bbox[0]
    y = joint_out[:, 1] / output_shape[0] * bbox[3] + bbox[1]
    z = (joint_out[:, 2] / depth_dim * 2. - 1.) * (bbox_3d_depth / 2.) + center_cam[2]
    return x, y, z


def pixel_to_cam(pixel_coord, f, c):
    x = (pixel_coord[..., 0] - c[0]) / f[0] * pixel_coord[..., 2]
    y = (pixel_coord[..., 1] - c[1]) / f[1] * pixel_coord[..., 2]
    z = pixel_coord[..., 2]
    return x, y, z


def evaluate_sample(pred, bbox, center_cam, f, c, root_idx, depth_dim, output_shape, bbox_3d_depth):
    """One sample (J, 3) float32 -> (pixel float32 (J,3), root-aligned camera float64 (J,3))."""
    pix = pred.copy()
    pix[:, 0], pix[:, 1], pix[:, 2] = warp_to_original(pix, bbox, center_cam, depth_dim, output_shape, bbox_3d_depth)
    cam = np.zeros((pred.shape[0], 3))
    cam[:, 0], cam[:, 1], cam[:, 2] = pixel_to_cam(pix, f, c)
    if root_idx is not None and root_idx >= 0:
        cam = cam - cam[root_idx]
    return pix, cam


def post_process(coords, flipped, flip_pairs, bbox, center_cam, f, c, root_idx, depth_dim, output_shape, bbox_3d_depth):
    """Batch driver: numpy in, (merged float32, pixel float32, cam float64) out, each (B, J, 3)."""
    merged = coords
    if flipped is not None:
        merged = flip_merge(torch.from_numpy(coords), torch.from_numpy(flipped), output_shape[1], flip_pairs).numpy()
    pix = np.zeros_like(merged)
    cam = np.zeros(merged.shape, np.float64)
    for n in range(merged.shape[0]):
        pix[n], cam[n] = evaluate_sample(merged[n], bbox[n], center_cam[n], f[n], c[n], root_idx, depth_dim, output_shape, bbox_3d_depth)
    return merged, pix, cam


def make_inputs(B, J, D, H, W, seed, flip):
    """Seeded inputs in Human3.6M-like ranges: boxes of a few hundred pixels, depth ~ 3-6 m, focal ~ 1145 px."""
    g = np.random.default_rng(seed)
    coords = (g.random((B, J, 3)) * [W - 1, H - 1, D - 1]).astype(np.float32)
    flipped = (g.random((B, J, 3)) * [W - 1, H - 1, D - 1]).astype(np.float32) if flip else None
    bbox = np.concatenate([g.random((B, 2)) * 600, 150 + g.random((B, 2)) * 500], 1).astype(np.float32)
    center = np.concatenate([g.normal(0, 500, (B, 2)), 3000 + g.random((B, 1)) * 3000], 1).astype(np.float32)
    f = (1145 + g.normal(0, 3, (B, 2))).astype(np.float32)
    c = (510 + g.normal(0, 8, (B, 2))).astype(np.float32)
    return coords, flipped, bbox, center, f, c
