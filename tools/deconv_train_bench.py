#!/usr/bin/env python
"""tools/deconv_train_bench.py -- the deconv block in TRAINING (row N1): ihpr_b200.deconv_bn_relu_train (K9 kTrain + K10 + K9 kDgrad) against
the stock module stack (cuDNN transposed convolution + BatchNorm2d in training mode + ReLU, bf16 channels_last), forward and
forward + backward.  CUDA events, median of --iters.  --profile: one forward + backward of the fused block only (for ncu)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ihpr_b200

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--H", type=int, default=32)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--variant", type=int, default=0, help="K11: 21 / 22 / 24 = clusters of 1 / 2 / 4 CTAs, 64-pixel stages; 31 / 32 / 34 = 32-pixel stages")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
ihpr_b200.set_variant(a.variant)
B, H = a.B, a.H
x = torch.randn(B, 256, H, H, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
deconv = torch.nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
torch.nn.init.normal_(deconv.weight, std=0.03)
bn = torch.nn.BatchNorm2d(256).to(dev)                                  # fp32 parameters, as under autocast
bn_bf = torch.nn.BatchNorm2d(256).to(dev).to(torch.bfloat16)
dout = torch.randn(B, 256, 2 * H, 2 * H, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def fused():
    return ihpr_b200.deconv_bn_relu_train(x, deconv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps)


def stock():
    return torch.relu(bn_bf(deconv(x)))


def step(fn):
    def run():
        x.grad = None
        deconv.weight.grad = None
        fn().backward(dout)
    return run


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return sorted(s.elapsed_time(e) for s, e in ev)[a.iters // 2] * 1e3


if a.profile:
    step(fused)()
    torch.cuda.synchronize()
    sys.exit(0)

with torch.no_grad():
    t_f, t_s = timeit(fused), timeit(stock)
t_fb, t_sb = timeit(step(fused)), timeit(step(stock))
step(fused)()
gx_f, gw_f = x.grad.float().clone(), deconv.weight.grad.float().clone()
step(stock)()
gx_s, gw_s = x.grad.float(), deconv.weight.grad.float()
flop = 2.0 * B * (2 * H) ** 2 * 256 * 1024
print(json.dumps({"B": B, "H": H, "wgrad": os.environ.get("IHPR_DECONV_WGRAD", "k11"), "variant": a.variant, "fused_fwd_us": round(t_f, 1), "stock_fwd_us": round(t_s, 1), "fused_fwd_bwd_us": round(t_fb, 1), "stock_fwd_bwd_us": round(t_sb, 1),
                  "fwd_speedup": round(t_s / t_f, 2), "fwd_bwd_speedup": round(t_sb / t_fb, 2), "gemm_GFLOP": round(flop / 1e9, 1),
                  "dx_rel_diff_vs_stock": round(float((gx_f - gx_s).norm() / gx_s.norm()), 5),
                  "dw_rel_diff_vs_stock": round(float((gw_f - gw_s).norm() / gw_s.norm()), 5)}))
