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
