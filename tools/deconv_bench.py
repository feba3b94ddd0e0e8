#!/usr/bin/env python
"""tools/deconv_bench.py -- K9 (deconv3 + BatchNorm + ReLU as one tcgen05 kernel) against the stock module stack (cuDNN transposed
convolution + BatchNorm + ReLU, bf16 channels_last, eval mode), and the inference head tail K9 -> K3 against stock -> conv1x1 -> K1.
CUDA events, median of --iters."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ihpr_b200

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--J", type=int, default=18)
ap.add_argument("--D", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variant", type=int, default=0, help="21 / 22 / 24: K9 on clusters of 1 / 2 / 4 CTAs")
a = ap.parse_args()
dev = torch.device("cuda:0")
ihpr_b200.set_variant(a.variant)
torch.backends.cudnn.benchmark = True
B, J, D = a.B, a.J, a.D
x = torch.randn(B, 256, 32, 32, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
deconv = torch.nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
torch.nn.init.normal_(deconv.weight, std=0.03)
bn = torch.nn.BatchNorm2d(256).to(dev)
bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
bn_bf = torch.nn.BatchNorm2d(256).to(dev).to(torch.bfloat16)
bn_bf.load_state_dict(bn.state_dict())
final = torch.nn.Conv2d(256, J * D, 1).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
torch.nn.init.normal_(final.weight, std=0.02)
for m in (deconv, bn, bn_bf, final):
    m.eval()


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return sorted(s.elapsed_time(e) for s, e in ev)[a.iters // 2] * 1e3


with torch.no_grad():
    k9 = lambda: ihpr_b200.deconv_bn_relu(x, deconv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)   # noqa: E731
    stock = lambda: torch.relu_(bn_bf(deconv(x)))                                                                       # noqa: E731
    t_k9, t_stock, t_conv = timeit(k9), timeit(stock), timeit(lambda: deconv(x))
    y9, ys = k9(), stock()
    wb, bias = final.weight.reshape(J * D, 256), final.bias.float()
    t_tail = timeit(lambda: ihpr_b200.fused_head_soft_argmax(k9(), wb, bias, J))
    t_tail_stock = timeit(lambda: ihpr_b200.soft_argmax(final(stock()), J))
    c9 = ihpr_b200.fused_head_soft_argmax(y9, wb, bias, J)
    cs = ihpr_b200.soft_argmax(final(ys), J)
flop = 2.0 * B * 64 * 64 * 256 * 1024
print(json.dumps({"variant": a.variant, "B": B, "k9_us": round(t_k9, 1), "k9_TFLOPs": round(flop / t_k9 / 1e6, 1), "stock_deconv_bn_relu_us": round(t_stock, 1),
                  "stock_deconv_only_us": round(t_conv, 1), "speedup": round(t_stock / t_k9, 2),
                  "max_rel_diff_vs_stock_bf16": round(((y9.float() - ys.float()).abs().max() / ys.float().abs().max()).item(), 5),
                  "tail_k9_k3_us": round(t_tail, 1), "tail_stock_conv_k1_us": round(t_tail_stock, 1), "tail_speedup": round(t_tail_stock / t_tail, 2),
                  "coords_max_diff": round((c9 - cs).abs().max().item(), 4)}))
