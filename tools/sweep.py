#!/usr/bin/env python
"""tools/sweep.py -- kernel sweep on one GPU (BASELINE.json configs[4]): variants x dtype x shape, CUDA-event timed.
Prints one JSON line per cell: fwd / bwd microseconds and achieved algorithmic GB/s (N*s and 2*N*s bytes per volume)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ihpr_b200
from ihpr_b200 import functional as F
from ihpr_b200._lib import check, lib


def time_cell(B, J, D, H, W, dtype, variant, iters, nbuf):
    dev = torch.device("cuda:0")
    es = 4 if dtype == torch.float32 else 2
    R, N = B * J, D * H * W
    ihpr_b200.set_variant(variant)
    heats = [torch.randn(B, J * D, H, W, device=dev).to(dtype) for _ in range(nbuf)]
    grads = [torch.empty_like(h) for h in heats]
    gt = torch.rand(B, J, 3, device=dev) * 64
    vis = torch.ones(B, J, device=dev)
    hd = torch.ones(B, 1, device=dev)
    go = torch.ones((), device=dev)
    L = lib()
    code = 0 if dtype == torch.float32 else 1
    stream = torch.cuda.current_stream().cuda_stream
    outs = []

    def fwd(i):
        coords, stats, loss = F._fwd(heats[i % nbuf], J, (gt, vis, hd))
        outs.append((coords, stats))
        return coords, stats

    def bwd(i, coords, stats):
        check(L.ihpr_integral_l1_bwd(heats[i % nbuf].data_ptr(), code, B, J, D, H, W, coords.data_ptr(), stats.data_ptr(), gt.data_ptr(),
                                     vis.data_ptr(), hd.data_ptr(), go.data_ptr(), grads[i % nbuf].data_ptr(), stream))

    for i in range(3):
        c, s = fwd(i); bwd(i, c, s)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3 * iters)]
    for i in range(iters):
        e[3 * i].record(); c, s = fwd(i); e[3 * i + 1].record(); bwd(i, c, s); e[3 * i + 2].record()
    torch.cuda.synchronize()
    f = sorted(e[3 * i].elapsed_time(e[3 * i + 1]) for i in range(iters))[iters // 2] * 1e3
    b = sorted(e[3 * i + 1].elapsed_time(e[3 * i + 2]) for i in range(iters))[iters // 2] * 1e3
    return {"B": B, "J": J, "D": D, "H": H, "W": W, "dtype": str(dtype).split(".")[-1], "variant": variant, "fwd_us": round(f, 1), "bwd_us": round(b, 1),
            "fwd_GBps": round(R * N * es / f / 1e3, 1), "bwd_GBps": round(2 * R * N * es / b / 1e3, 1),
            "fwdbwd_GBps": round(3 * R * N * es / (f + b) / 1e3, 1), "vol_per_s": round(R / ((f + b) * 1e-6))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="variants", choices=["variants", "shapes"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    if args.mode == "variants":
        for dtype in (torch.float32, torch.bfloat16):
            for v in (0, 1, 11, 12, 13, 14, 2, 21):
                print(json.dumps(time_cell(32, 18, 64, 64, 64, dtype, v, args.iters, 1)), flush=True)
    else:
        for dtype in (torch.float32, torch.bfloat16):
            for D in (32, 64, 128):
                for J in (17, 18):
                    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
                        if B * J * D * 64 * 64 * 4 * 2 > 40 << 30:
                            continue
                        # small shapes: rotate enough buffers that consecutive iterations do not hit in L2
                        vol = B * J * D * 64 * 64 * (4 if dtype == torch.float32 else 2)
                        nbuf = max(1, min(8, (512 << 20) // max(vol, 1)))
                        print(json.dumps(time_cell(B, J, D, 64, 64, dtype, 0, args.iters, nbuf)), flush=True)


if __name__ == "__main__":
    main()
