"""oracle/inputs.py -- TEST INFRASTRUCTURE: deterministic synthetic heatmaps / targets.

Uses numpy's frozen legacy ``RandomState`` stream so the same (dist, shape, seed) gives the same
bytes in the build container (where goldens are produced from the reference) and on the GPU box
(where the CUDA path is compared with them).  Distributions follow SURVEY.md 8c / BASELINE.md 5.
"""
import numpy as np

DISTS = ("randn1", "randn3", "init", "blobs", "large", "shifted")


def make_heat(dist, B, J, D, H, W, seed):
    rs = np.random.RandomState(seed)
    shape = (B, J * D, H, W)
    if dist == "randn1":
        h = rs.standard_normal(shape)
    elif dist == "randn3":
        h = 3.0 * rs.standard_normal(shape)
    elif dist == "init":            # what a freshly initialised final_layer emits (model.py:53-56)
        h = 1e-3 * rs.standard_normal(shape)
    elif dist == "large":           # |h| >> 88: only finite thanks to max-subtraction
        h = 50.0 * rs.standard_normal(shape) + 300.0
    elif dist == "shifted":
        h = rs.standard_normal(shape) - 1000.0
    elif dist == "blobs":           # peaked Gaussian blob per joint, amp 20, sigma 2 voxels, + 0.1 noise
        z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
        h = np.empty((B, J, D, H, W))
        for b in range(B):
            for j in range(J):
                c = rs.uniform(0, 1, 3) * np.array([W - 1, H - 1, D - 1])
                d2 = (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2
                h[b, j] = 20.0 * np.exp(-d2 / (2 * 2.0 ** 2))
        h = h.reshape(shape) + 0.1 * rs.standard_normal(shape)
    else:
        raise ValueError(dist)
    return np.ascontiguousarray(h, dtype=np.float32)


def make_targets(B, J, D, H, W, seed, vis_mode="ones", hd_mode="ones"):
    """gt_coord (B,J,3) in voxel units (x<W, y<H, z<D), gt_vis (B,J,1), gt_have_depth (B,1):
    shapes as produced by /root/reference/data/dataset.py:146-152."""
    rs = np.random.RandomState(seed + 7919)
    gt = rs.uniform(0, 1, (B, J, 3)) * np.array([W, H, D])
    vis = np.ones((B, J, 1)) if vis_mode == "ones" else (rs.uniform(0, 1, (B, J, 1)) > 0.4).astype(np.float64)
    if hd_mode == "ones":
        hd = np.ones((B, 1))
    elif hd_mode == "zeros":
        hd = np.zeros((B, 1))
    else:
        hd = (np.arange(B).reshape(B, 1) % 2).astype(np.float64)
    return gt.astype(np.float32), vis.astype(np.float32), hd.astype(np.float32)
