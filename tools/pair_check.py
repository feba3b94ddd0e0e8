#!/usr/bin/env python
"""tools/pair_check.py -- the SM-pair (cta_group::2) form of K3 / K4 (variant 5) against the single-SM form (variant 0) on the same inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ihpr_b200
from ihpr_b200._lib import lib, check

dev = torch.device("cuda:0")
ok = True
for (B, J, D, K, H, W) in [(2, 18, 64, 256, 64, 64), (3, 17, 64, 256, 64, 64), (5, 5, 32, 128, 32, 32), (1, 3, 128, 192, 16, 32), (40, 2, 32, 64, 8, 32)]:
    g = torch.Generator(device="cpu").manual_seed(B)
    x = torch.randn(B, K, H, W, generator=g).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)
    wb = (torch.randn(J * D, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = (torch.randn(J * D, generator=g) * 0.5).to(dev)
    gt = (torch.rand(B, J, 3, generator=g) * torch.tensor([W, H, D], dtype=torch.float32)).to(dev)
    vis, hd, go = torch.ones(B, J, device=dev), torch.ones(B, 1, device=dev), torch.full((), 1.5, device=dev)
    res = {}
    for v in (0, 5):
        ihpr_b200.set_variant(v)
        with torch.no_grad():
            coords, stats = ihpr_b200.functional.fused_head_soft_argmax(x, wb, bias, J, return_stats=True)
        dheat = torch.full((B, J * D, H * W), float("nan"), dtype=torch.bfloat16, device=dev)
        dbp = torch.full((B, 4, J * D), float("nan"), device=dev)
        check(lib().ihpr_head_integral_l1_bwd(x.data_ptr(), wb.data_ptr(), bias.data_ptr(), B, K, J, D, H, W, coords.data_ptr(), stats.data_ptr(),
                                              gt.data_ptr(), vis.data_ptr(), hd.data_ptr(), go.data_ptr(), dheat.data_ptr(), dbp.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        res[v] = (coords.clone(), stats.clone(), dheat.float(), dbp.sum((0, 1)))
    ihpr_b200.set_variant(0)
    dc = (res[0][0] - res[5][0]).abs().max().item()
    ds = (res[0][1] - res[5][1]).abs().max().item()
    dg = (res[0][2] - res[5][2]).abs().max().item() / res[0][2].abs().max().item()
    db = (res[0][3] - res[5][3]).abs().max().item() / res[0][3].abs().max().item()
    nan = bool(torch.isnan(res[5][2]).any() or torch.isnan(res[5][3]).any() or torch.isnan(res[5][0]).any())
    good = dc <= 1e-4 and ds <= 1e-3 and dg <= 1e-2 and db <= 1e-4 and not nan
    ok = ok and good
    print("B=%d J=%d D=%d K=%d %dx%d: |dcoords| %.2e |dstats| %.2e grad rel %.2e dbias rel %.2e nan %s -> %s" % (B, J, D, K, H, W, dc, ds, dg, db, nan, "ok" if good else "MISMATCH"))
sys.exit(0 if ok else 1)
