"""Timing of the device-side sample preparation (K7 patches + K8 joints) against the reference's per-sample CPU recipe
(cv2.warpAffine + numpy tail, as data/dataset.py:84-152 runs it in its DataLoader workers).  Usage: python tools/aug_bench.py [B] [side]"""
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ihpr_b200                      # noqa: E402
from ihpr_b200 import data            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
side = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
J = 18
g = np.random.default_rng(0)
imgs = g.integers(0, 256, (B, side, side, 3), dtype=np.uint8)
bbox = np.stack([np.array([side * 0.2, side * 0.1, side * 0.5, side * 0.7], np.float32)] * B)
joints = np.concatenate([g.random((B, J, 2)) * side, g.normal(0, 400, (B, J, 1))], 2)
vis = np.ones((B, J))
np.random.seed(0); random.seed(0)
augs = [data.get_aug_config() for _ in range(B)]
pairs = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))
dev = torch.device("cuda:0")
pinned = torch.from_numpy(imgs).pin_memory()
d_imgs = pinned.to(dev)


def step(images):
    return ihpr_b200.augment_batch(images, [[side, side]] * B, bbox, joints, vis, augs, flip_pairs=pairs)


for _ in range(3):
    step(d_imgs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters):
    step(d_imgs)
e1.record(); torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) / iters
t0 = time.perf_counter()
for _ in range(iters):
    out = step(pinned.to(dev, non_blocking=True))
torch.cuda.synchronize()
t_e2e = (time.perf_counter() - t0) / iters * 1e3
print("K7+K8 B=%d %dx%d -> 256x256: images resident %.3f ms/batch (%.1f k samples/s); with H2D of the raw images %.3f ms/batch (%.1f k samples/s, %.1f MB/batch)"
      % (B, side, side, t_dev, B / t_dev, t_e2e, B / t_e2e, imgs.nbytes / 1e6))
try:
    import cv2
    cv2.setNumThreads(1)
    mean, std = np.array(data.PIXEL_MEAN, np.float32).reshape(3, 1, 1), np.array(data.PIXEL_STD, np.float32).reshape(3, 1, 1)
    t0 = time.perf_counter()
    for b in range(B):
        scale, rot, flip, cs = augs[b]
        img = imgs[b][:, ::-1] if flip else imgs[b]
        tr = data.patch_params(bbox[b], side, augs[b], (256, 256))
        p = cv2.warpAffine(np.ascontiguousarray(img), tr, (256, 256), flags=cv2.INTER_LINEAR)[:, :, ::-1].astype(np.float32)
        for i in range(3):
            p[:, :, i] = np.clip(p[:, :, i] * cs[i], 0, 255)
        o = (np.ascontiguousarray(p.transpose(2, 0, 1)) - mean) / std
    t_cpu = (time.perf_counter() - t0) / B * 1e3
    print("reference recipe on one host core (cv2 %s): %.3f ms/sample (%.2f k samples/s per core)" % (cv2.__version__, t_cpu, 1 / t_cpu))
except ImportError:
    print("cv2 not available: no CPU figure")
