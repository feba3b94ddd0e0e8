#!/usr/bin/env python
"""tools/head_bench.py -- K3 (fused 1x1 conv + soft-argmax, tcgen05) vs the unfused tail (cuDNN/cuBLAS conv in bf16
channels_last writing the heat-map, then K1 reading it).  CUDA events, median of `iters`."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ihpr_b200

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--J", type=int, default=18)
ap.add_argument("--D", type=int, default=64)
ap.add_argument("--K", type=int, default=256)
ap.add_argument("--hw", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variant", type=int, default=0, help="3 = K4w / K4x on SM pairs (tcgen05 cta_group::2)")
a = ap.parse_args()
dev = torch.device("cuda:0")
ihpr_b200.set_variant(a.variant)
B, J, D, K, H, W = a.B, a.J, a.D, a.K, a.hw, a.hw
x = torch.randn(B, K, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
conv = torch.nn.Conv2d(K, J * D, 1).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
torch.nn.init.normal_(conv.weight, std=0.05)
wt, bias = conv.weight.detach(), conv.bias.detach().float()
torch.backends.cudnn.benchmark = True


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
    wb = wt.reshape(J * D, K).contiguous()
    t_fused = timeit(lambda: ihpr_b200.functional.fused_head_soft_argmax(x, wb, bias, J))
    t_conv = timeit(lambda: conv(x))
    heat = conv(x)
    t_k1 = timeit(lambda: ihpr_b200.soft_argmax(heat, J))
    t_unf = timeit(lambda: ihpr_b200.soft_argmax(conv(x), J))
    c1 = ihpr_b200.functional.fused_head_soft_argmax(x, wb, bias, J)
# K4: heat-map gradient straight from the TMEM recompute
from ihpr_b200._lib import lib, check
gt = torch.rand(B, J, 3, device=dev) * torch.tensor([W, H, D], device=dev, dtype=torch.float32)
vis, hd = torch.ones(B, J, device=dev), torch.ones(B, 1, device=dev)
with torch.no_grad():
    coords, stats = ihpr_b200.functional.fused_head_soft_argmax(x, wb, bias, J, return_stats=True)
go = torch.ones((), device=dev)
dheat = torch.empty(B, J * D, H * W, dtype=torch.bfloat16, device=dev)
stream = torch.cuda.current_stream().cuda_stream
t_k4 = timeit(lambda: check(lib().ihpr_head_integral_l1_bwd(x.data_ptr(), wb.data_ptr(), bias.data_ptr(), B, K, J, D, H, W, coords.data_ptr(),
                                                            stats.data_ptr(), gt.data_ptr(), vis.data_ptr(), hd.data_ptr(), go.data_ptr(),
                                                            dheat.data_ptr(), None, stream)))
# K4w / K4x: dW, dbias, dX in-kernel (no heat-map gradient in HBM, no library GEMM), and the library-GEMM comparison arm on K4's output
L = lib()
nbytes = L.ihpr_head_bwd_workspace_bytes(B, K, J, D, H, W)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dx = torch.empty(B, H, W, K, dtype=torch.bfloat16, device=dev)
dw = torch.empty(J * D, K, device=dev)
db = torch.empty(J * D, device=dev)


def params(dxp, dwp, dbp):
    check(L.ihpr_head_integral_l1_bwd_params(x.data_ptr(), wb.data_ptr(), bias.data_ptr(), B, K, J, D, H, W, coords.data_ptr(), stats.data_ptr(),
                                             gt.data_ptr(), vis.data_ptr(), hd.data_ptr(), go.data_ptr(), dxp, dwp, dbp, ws.data_ptr(), nbytes, stream))


t_k4w = timeit(lambda: params(None, dw.data_ptr(), db.data_ptr()))
t_k4x = timeit(lambda: params(dx.data_ptr(), None, None))
t_k4wx = timeit(lambda: params(dx.data_ptr(), dw.data_ptr(), db.data_ptr()))
xn = x.permute(0, 2, 3, 1).reshape(B, H * W, K)
with torch.no_grad():
    t_lib_dx = timeit(lambda: torch.matmul(dheat.transpose(1, 2), wb))
    t_lib_dw = timeit(lambda: torch.bmm(dheat, xn, out_dtype=torch.float32).sum(0))
with torch.no_grad():
    c2 = ihpr_b200.soft_argmax(conv(x).float(), J)
flop = 2.0 * B * J * D * K * H * W
print(json.dumps({"B": B, "k4w_dW_us": round(t_k4w, 1), "k4w_TFLOPs_2gemm": round(2 * flop / t_k4w / 1e6, 1), "k4x_dX_us": round(t_k4x, 1),
                  "k4x_TFLOPs_2gemm": round(2 * flop / t_k4x / 1e6, 1), "k4w_plus_k4x_us": round(t_k4wx, 1),
                  "algorithmic_TFLOPs_dW_dX_only": round(2 * flop / t_k4wx / 1e6, 1), "executed_TFLOPs_4gemm": round(4 * flop / t_k4wx / 1e6, 1),
                  "comparison_arm": {"k4_grad_heat_us": round(t_k4, 1), "library_dx_us": round(t_lib_dx, 1), "library_dw_us": round(t_lib_dw, 1),
                                     "total_us": round(t_k4 + t_lib_dx + t_lib_dw, 1)}}))
print(json.dumps({"variant": a.variant, "B": B, "J": J, "D": D, "K": K, "HW": H, "fused_us": round(t_fused, 1), "fused_TFLOPs": round(flop / t_fused / 1e6, 1),
                  "k4_bwd_us": round(t_k4, 1), "k4_TFLOPs": round(flop / t_k4 / 1e6, 1), "k4_write_GBps": round(B * J * D * H * W * 2 / t_k4 / 1e3, 1),
                  "conv_us": round(t_conv, 1), "k1_bf16_us": round(t_k1, 1), "conv_plus_k1_us": round(t_unf, 1),
                  "speedup": round(t_unf / t_fused, 2), "max_coord_diff_vs_unfused": float((c1 - c2).abs().max())}))
