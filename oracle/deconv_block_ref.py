"""oracle/deconv_block_ref.py -- TEST INFRASTRUCTURE (checker only; never imported by the product path).

numpy fp64 restatement of ONE block of the reference's ``HeadNet._make_deconv_layer`` (/root/reference/main/model.py:22-38):

    nn.ConvTranspose2d(C_in, C_out, kernel_size=4, stride=2, padding=1, output_padding=0, bias=False)      model.py:25-33
    nn.BatchNorm2d(C_out)                                                                                   model.py:34
    nn.ReLU(inplace=True)                                                                                   model.py:35

in training mode (batch statistics, as under main/train.py:64-71) and in eval mode (running statistics, main/test.py:62), forward and
backward.  The three modules are PyTorch's (torch 1.0 in the reference's requirements, 2.11 in this image -- absent from /root/reference);
their published definitions are restated here:

  * ConvTranspose2d (torch.nn docs, "fractionally-strided convolution"): every input pixel (iy, ix) adds x[b, ci, iy, ix] * w[ci, co, ky, kx]
    to output pixel (oy, ox) = (iy * stride - padding + ky, ix * stride - padding + kx); weight layout (C_in, C_out, kH, kW).
  * BatchNorm2d, training: y = (x - E[x]) / sqrt(Var[x] + eps) * gamma + beta with the BIASED variance over (N, H, W); the running buffers
    are updated as running = (1 - momentum) * running + momentum * batch with the UNBIASED variance (momentum 0.1, eps 1e-5 by default).
  * ReLU: max(0, x).

Pinned against the reference's own modules (oracle/make_golden.py --deconv builds main/model.py's HeadNet and runs its third block in
fp64; tests/golden/deconv_block_train_*.npz) by tests/test_oracle.py.
"""
import numpy as np


def problem(seed, B, Cin, Cout, H, W):
    """Seeded inputs shared by the golden generator and the tests (values exactly representable in bf16, so that every
    implementation sees the same operands): x, weight (C_in, C_out, 4, 4), gamma, beta, running_mean, running_var, dout."""
    r = np.random.RandomState(seed)

    def bf16(a):
        u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        u = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)        # round to nearest even, keep 8 mantissa bits
        return u.view(np.float32).astype(np.float64)

    x = bf16(r.randn(B, Cin, H, W))
    w = bf16(r.randn(Cin, Cout, 4, 4) * 0.05)
    gamma = (r.rand(Cout) + 0.5).astype(np.float32).astype(np.float64)
    beta = (r.randn(Cout) * 0.3).astype(np.float32).astype(np.float64)
    rm = (r.randn(Cout) * 0.1).astype(np.float32).astype(np.float64)
    rv = (r.rand(Cout) + 0.5).astype(np.float32).astype(np.float64)
    dout = bf16(r.randn(B, Cout, 2 * H, 2 * W))
    return x, w, gamma, beta, rm, rv, dout


def _taps(H, W):
    """for every kernel tap (ky, kx): the input window [iy0:iy1, ix0:ix1] whose output pixels 2*iy - 1 + ky, 2*ix - 1 + kx are in range"""
    out = []
    for ky in range(4):
        iy0, iy1 = (1 if ky == 0 else 0), (H - 1 if ky == 3 else H)
        for kx in range(4):
            ix0, ix1 = (1 if kx == 0 else 0), (W - 1 if kx == 3 else W)
            out.append((ky, kx, iy0, iy1, ix0, ix1))
    return out


def conv_transpose2d(x, w):
    """model.py:25-33: kernel 4, stride 2, padding 1, no bias.  x (B, C_in, H, W), w (C_in, C_out, 4, 4) -> (B, C_out, 2H, 2W)"""
    B, Cin, H, W = x.shape
    y = np.zeros((B, w.shape[1], 2 * H, 2 * W), dtype=np.float64)
    for ky, kx, iy0, iy1, ix0, ix1 in _taps(H, W):
        contrib = np.einsum("bchw,cd->bdhw", x[:, :, iy0:iy1, ix0:ix1], w[:, :, ky, kx])
        y[:, :, 2 * iy0 - 1 + ky:2 * iy1 - 1 + ky:2, 2 * ix0 - 1 + kx:2 * ix1 - 1 + kx:2] += contrib
    return y


def forward_train(x, w, gamma, beta, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    """the block in training mode: returns dict(y = raw convolution output, out, mean, var (biased), running_mean, running_var (updated copies))"""
    y = conv_transpose2d(x, w)
    n = y.shape[0] * y.shape[2] * y.shape[3]
    mean = y.mean(axis=(0, 2, 3))
    var = y.var(axis=(0, 2, 3))                                      # biased, what normalises
    xhat = (y - mean[None, :, None, None]) / np.sqrt(var + eps)[None, :, None, None]
    out = np.maximum(xhat * gamma[None, :, None, None] + beta[None, :, None, None], 0.0)
    res = {"y": y, "out": out, "mean": mean, "var": var}
    if running_mean is not None:
        res["running_mean"] = (1 - momentum) * running_mean + momentum * mean
        res["running_var"] = (1 - momentum) * running_var + momentum * var * n / (n - 1)
    return res


def forward_eval(x, w, gamma, beta, running_mean, running_var, eps=1e-5):
    """the block in eval mode (main/test.py:62): running statistics"""
    y = conv_transpose2d(x, w)
    s = gamma / np.sqrt(running_var + eps)
    return np.maximum(y * s[None, :, None, None] + (beta - running_mean * s)[None, :, None, None], 0.0)


def backward_train(x, w, gamma, beta, dout, eps=1e-5):
    """gradients of sum(out * dout) for the training-mode block: dict(dx, dw, dgamma, dbeta, dy = gradient of the raw convolution output)"""
    f = forward_train(x, w, gamma, beta, eps=eps)
    y, mean, var = f["y"], f["mean"], f["var"]
    n = y.shape[0] * y.shape[2] * y.shape[3]
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (y - mean[None, :, None, None]) * rstd[None, :, None, None]
    dz = dout * (f["out"] > 0)                                       # ReLU
    dbeta = dz.sum(axis=(0, 2, 3))
    dgamma = (dz * xhat).sum(axis=(0, 2, 3))
    dy = (gamma * rstd)[None, :, None, None] * (dz - dbeta[None, :, None, None] / n - xhat * dgamma[None, :, None, None] / n)
    B, Cin, H, W = x.shape
    dx = np.zeros_like(x)
    dw = np.zeros_like(w)
    for ky, kx, iy0, iy1, ix0, ix1 in _taps(H, W):
        g = dy[:, :, 2 * iy0 - 1 + ky:2 * iy1 - 1 + ky:2, 2 * ix0 - 1 + kx:2 * ix1 - 1 + kx:2]     # the output pixels this tap wrote
        dx[:, :, iy0:iy1, ix0:ix1] += np.einsum("bdhw,cd->bchw", g, w[:, :, ky, kx])
        dw[:, :, ky, kx] = np.einsum("bchw,bdhw->cd", x[:, :, iy0:iy1, ix0:ix1], g)
    return {"dx": dx, "dw": dw, "dgamma": dgamma, "dbeta": dbeta, "dy": dy}
