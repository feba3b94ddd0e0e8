#!/usr/bin/env python
"""tools/sanitize_all.py -- one small invocation of every kernel family of the library (K1/K2 ring, direct and scalar paths, K1c, K5, K5c with
clusters of 2 / 8 / 16, K3 / K4w / K4x single-SM and SM-pair, K6, K9 at inference and in training with K10 / K11) with finiteness checks: a quick all-kernel smoke on a B200, and the
driver to put under `compute-sanitizer --tool memcheck|racecheck|synccheck` where the pool allows it (round 1's pool does not)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ihpr_b200
from ihpr_b200._lib import lib, check

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def targets(B, J, D, H, W):
    gt = (torch.rand(B, J, 3, generator=g) * torch.tensor([W, H, D], dtype=torch.float32)).to(dev)
    return gt, torch.ones(B, J, 1, device=dev), torch.ones(B, 1, device=dev)


def loss_step(B, J, D, H, W, dtype, fused, variant=0):
    ihpr_b200.set_variant(variant)
    h = torch.randn(B, J * D, H, W, generator=g).to(dev).to(dtype).requires_grad_(True)
    gt, vis, hd = targets(B, J, D, H, W)
    loss = ihpr_b200.integral_l1_loss(h, gt, vis, hd, fused_backward=fused)
    (loss * 1.25).backward()
    torch.cuda.synchronize()
    ihpr_b200.set_variant(0)
    assert torch.isfinite(loss).item() and torch.isfinite(h.grad.float()).all().item()


# K1 + K2 (ring, direct, scalar paths), fp32 and bf16
for v in (0, 2):
    loss_step(2, 3, 8, 16, 16, torch.float32, False, v)
loss_step(2, 3, 8, 16, 16, torch.bfloat16, False)
loss_step(2, 2, 3, 5, 9, torch.float32, False)                     # scalar path
loss_step(1, 18, 64, 64, 64, torch.float32, False)                 # K1c: clusters of 4 with a DSMEM merge (small batch)
loss_step(4, 17, 64, 64, 64, torch.bfloat16, False)                # K1c: clusters of 2 (bf16)
loss_step(2, 3, 8, 16, 16, torch.float32, False, 11)               # the persistent ring kernel on a small batch
# K5 (cooperative, L2-resident) and K5c (cluster-resident), enough joint-volumes for the one-launch path
loss_step(20, 16, 8, 16, 16, torch.float32, True)                  # 8 KiB volumes: S = 1
loss_step(10, 18, 32, 32, 32, torch.float32, True)                 # 128 KiB volumes
loss_step(10, 18, 32, 32, 32, torch.float32, True, variant=7)      # K5c, clusters of 2
loss_step(4, 18, 64, 64, 64, torch.float32, True, variant=7)       # K5c, clusters of 16
loss_step(6, 18, 64, 64, 64, torch.bfloat16, True, variant=7)      # K5c, clusters of 8
loss_step(6, 18, 64, 64, 64, torch.float32, True)                  # K5, S = 12
# K3 forward + K4w / K4x backward (one CTA per SM, and both backward kernels on SM pairs)
for v in (0, 3):
    ihpr_b200.set_variant(v)
    B, J, D, K, H, W = 3, 5, 32, 128, 32, 32
    x = torch.randn(B, K, H, W, generator=g).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wt = (torch.randn(J * D, K, 1, 1, generator=g) * 0.05).to(torch.bfloat16).to(dev).requires_grad_(True)
    bias = (torch.randn(J * D, generator=g) * 0.5).to(dev).requires_grad_(True)
    gt, vis, hd = targets(B, J, D, H, W)
    loss = ihpr_b200.fused_head_integral_l1_loss(x, wt, bias, gt, vis, hd)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item() and torch.isfinite(x.grad.float()).all().item()
ihpr_b200.set_variant(0)
# K9: deconv3 + BatchNorm + ReLU (full items and half items)
for Bq in (1, 10):
    xq = torch.randn(Bq, 256, 32, 32, generator=g).to(torch.bfloat16).to(dev)
    wq = (torch.randn(256, 256, 4, 4, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    with torch.no_grad():
        yq = ihpr_b200.deconv_bn_relu(xq, wq, torch.ones(256, device=dev), torch.zeros(256, device=dev), torch.zeros(256, device=dev),
                                      torch.ones(256, device=dev), 1e-5)
    torch.cuda.synchronize()
    assert torch.isfinite(yq.float()).all().item()
# K9 kTrain + K10 + K9 kDgrad + K11: the deconv block in training, forward and backward (32- and 16-wide maps; K11 also on clusters of 2 / 4)
for (Bq, Hq, vq) in ((2, 32, 0), (3, 16, 0), (2, 32, 22), (2, 32, 24), (2, 32, 31)):
    ihpr_b200.set_variant(vq)
    xq = torch.randn(Bq, 256, Hq, Hq, generator=g).to(dev).requires_grad_(True)
    wq = (torch.randn(256, 256, 4, 4, generator=g) * 0.05).to(dev).requires_grad_(True)
    gq, bq = torch.ones(256, device=dev, requires_grad=True), torch.zeros(256, device=dev, requires_grad=True)
    rm, rv = torch.zeros(256, device=dev), torch.ones(256, device=dev)
    yq = ihpr_b200.deconv_bn_relu_train(xq, wq, gq, bq, rm, rv, 0.1, 1e-5)
    yq.float().square().sum().backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(t.float()).all().item() for t in (yq, xq.grad, wq.grad, gq.grad, bq.grad, rm, rv))
ihpr_b200.set_variant(0)
# K6: test-time post-processing
B, J = 3, 18
coords = torch.rand(B, J, 3, device=dev) * 64
flipped = torch.rand(B, J, 3, device=dev) * 64
out = ihpr_b200.coords_to_camera(coords, bbox=torch.tensor([[10., 20., 200., 220.]] * B, device=dev),
                                 center_cam=torch.tensor([[0., 0., 5000.]] * B, device=dev), f=torch.tensor([[1145., 1143.]] * B, device=dev),
                                 c=torch.tensor([[512., 515.]] * B, device=dev), root_idx=0, flipped_coord_out=flipped,
                                 flip_pairs=((1, 4), (2, 5), (3, 6)))
torch.cuda.synchronize()
print("sanitize_all: every kernel family ran once")
