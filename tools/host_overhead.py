#!/usr/bin/env python
"""tools/host_overhead.py -- CPU cost per JointLocationLoss fwd+bwd call (tiny shape, so GPU time is negligible)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ihpr_b200
dev = torch.device("cuda:0")
B, J, D, H, W = 1, 2, 4, 4, 4
h = torch.randn(B, J * D, H, W, device=dev, requires_grad=True)
gt, vis, hd = torch.rand(B, J, 3, device=dev), torch.ones(B, J, 1, device=dev), torch.ones(B, 1, device=dev)
crit = ihpr_b200.JointLocationLoss()
for _ in range(50):
    h.grad = None; crit(h, gt, vis, hd).backward()
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    loss = crit(h, gt, vis, hd)
t1 = time.perf_counter()
for _ in range(n):
    h.grad = None; crit(h, gt, vis, hd).backward()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host us/call: forward %.1f, forward+backward %.1f" % ((t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6))
# inference call (main/test.py:62-65 at cfg.test_batch_size = 4) and the one-launch step
with torch.no_grad():
    for B2 in (1, 4):
        h2 = torch.randn(B2, 18 * 64, 64, 64, device=dev)
        for _ in range(20):
            ihpr_b200.soft_argmax(h2, 18)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(500):
            ihpr_b200.soft_argmax(h2, 18)
        t_issue = (time.perf_counter() - t0) / 500 * 1e6
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            ihpr_b200.soft_argmax(h2, 18)
        e1.record()
        torch.cuda.synchronize()
        print("soft_argmax B=%d J=18 64^3 under no_grad: host issue %.1f us/call, device-paced %.1f us/call" % (B2, t_issue, e0.elapsed_time(e1) / 500 * 1e3))
t0 = time.perf_counter()
for _ in range(n):
    h.grad = None; crit.forward_backward(h, gt, vis, hd)
torch.cuda.synchronize()
print("host us/call: forward_backward (one call) %.1f" % ((time.perf_counter() - t0) / n * 1e6))
