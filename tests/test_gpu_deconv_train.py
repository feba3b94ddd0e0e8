"""GPU: the deconv block of HeadNet in TRAINING (main/model.py:22-38 under main/train.py:64-71) --
ConvTranspose2d(256, 256, k4 s2 p1) + BatchNorm2d with batch statistics + ReLU, forward and backward, through the C-ABI
(ihpr_deconv_bn_relu_train_fwd / _bwd: K9 in its kTrain / kDgrad modes, K10 BatchNorm passes).

Truth: torch's own conv_transpose2d / batch_norm(training=True) / relu and autograd in fp64 on the same bf16-rounded operands.
Bounds (stated where used): the raw convolution output is rounded to bf16 once (2^-8 relative); the batch statistics are those of the
ROUNDED output, so mean / rstd agree with the fp64 truth to ~1e-3 of a standard deviation; everything downstream is bf16 again.
Each stage is also checked TIGHTLY against fp64 arithmetic on the kernel's own upstream output (so that a wrong tap, a wrong constant
or a lost partial cannot hide behind the loose end-to-end bound)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

EPS = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _problem(B, Hin, Win, seed, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 256, Hin, Win, generator=g).to(torch.bfloat16)
    w = (torch.randn(256, 256, 4, 4, generator=g) * 0.05).to(torch.bfloat16)
    gamma = torch.rand(256, generator=g) + 0.5
    beta = torch.randn(256, generator=g) * 0.3
    return x.to(dev), w.to(dev), gamma.to(dev), beta.to(dev)


def _raw_calls(x, w, gamma, beta, rm=None, rv=None, momentum=0.1, dout=None, want_dx=True):
    """the two C-ABI entries, called directly: returns y_raw, out, saved[, dy_raw, dgamma, dbeta, dx] (NCHW-shaped channels_last tensors)"""
    from ihpr_b200._lib import lib, check
    L = lib()
    dev = x.device
    B, Cin, H, W = x.shape
    xb = x.contiguous(memory_format=torch.channels_last)
    y_raw = torch.empty((B, 256, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    out = torch.empty_like(y_raw)
    saved = torch.empty((4, 256), dtype=torch.float32, device=dev)
    n = L.ihpr_deconv_train_workspace_bytes(256, 256)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(L.ihpr_deconv_bn_relu_train_fwd(xb.data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr() if rm is not None else None,
                                          rv.data_ptr() if rv is not None else None, momentum, EPS, B, 256, 256, H, W, y_raw.data_ptr(), out.data_ptr(),
                                          saved.data_ptr(), ws.data_ptr(), n, stream))
    if dout is None:
        return y_raw, out, saved
    go = dout.contiguous(memory_format=torch.channels_last)
    dy = torch.empty_like(y_raw)
    dgb = torch.empty((2, 256), dtype=torch.float32, device=dev)
    dx = torch.empty((B, 256, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last) if want_dx else None
    check(L.ihpr_deconv_bn_relu_train_bwd(go.data_ptr(), y_raw.data_ptr(), w.data_ptr(), saved.data_ptr(), B, 256, 256, H, W, dy.data_ptr(), dgb[0].data_ptr(),
                                          dgb[1].data_ptr(), dx.data_ptr() if dx is not None else None, ws.data_ptr(), n, stream))
    return y_raw, out, saved, dy, dgb[0], dgb[1], dx


CASES = [
    (2, 32, 32),        # the head's third block: 32 x 32 -> 64 x 64
    (3, 8, 32),         # one 8-row group per sample: every tile touches the top AND the bottom border
    (5, 16, 16),        # the head's second block: 16 x 16 -> 32 x 32
    (10, 32, 32),       # 160 forward items on 148 SMs: the last wave runs as half items; CTAs own several items
    (40, 32, 32),       # 160 input-gradient items: half items in the kDgrad mode too
    (2, 32, 16),        # 16 wide, 32 high: two 16-row groups per sample
]


@pytest.mark.parametrize("case", CASES)
def test_training_forward_vs_torch_fp64(case, dev):
    B, Hin, Win = case
    x, w, gamma, beta = _problem(B, Hin, Win, seed=B * 100 + Hin + Win, dev=dev)
    rm0, rv0 = torch.randn(256, device=dev) * 0.1, torch.rand(256, device=dev) + 0.5
    rm, rv = rm0.clone(), rv0.clone()
    y_raw, out, saved = _raw_calls(x, w, gamma, beta, rm, rv, momentum=0.1)
    torch.cuda.synchronize()
    y64 = F.conv_transpose2d(x.double(), w.double(), stride=2, padding=1)
    # (1) the GEMM: one bf16 rounding of an fp32 accumulation of 1024 products
    err = (y_raw.double() - y64).abs()
    assert bool((err <= 2.0 ** -8 * y64.abs() + 2e-3 * y64.abs().max()).all()), float(err.max())
    # (2) the statistics, tightly: fp64 mean / biased variance of the kernel's OWN bf16 output
    n = y64.numel() // 256
    m_own = y_raw.double().mean(dim=(0, 2, 3))
    v_own = y_raw.double().var(dim=(0, 2, 3), unbiased=False)
    mean, rstd, scale, shift = (saved[i].double() for i in range(4))
    assert float((mean - m_own).abs().max()) <= 1e-5 * float(v_own.sqrt().max()) + 1e-6
    np.testing.assert_allclose(rstd.cpu().numpy(), (1.0 / torch.sqrt(v_own + EPS)).cpu().numpy(), rtol=2e-5)
    np.testing.assert_allclose(scale.cpu().numpy(), (gamma.double() * rstd).cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(shift.cpu().numpy(), (beta.double() - mean * scale).cpu().numpy(), rtol=1e-5, atol=1e-6)
    # ... and against the fp64 truth: the rounding noise of ~n values averages out
    m64, v64 = y64.mean(dim=(0, 2, 3)), y64.var(dim=(0, 2, 3), unbiased=False)
    assert float(((mean - m64).abs() / v64.sqrt()).max()) <= 1e-3
    np.testing.assert_allclose(rstd.cpu().numpy(), (1.0 / torch.sqrt(v64 + EPS)).cpu().numpy(), rtol=1e-3)
    # (3) running statistics exactly as torch.nn.BatchNorm2d updates them (momentum 0.1, unbiased variance)
    np.testing.assert_allclose(rm.cpu().numpy(), (0.9 * rm0.double() + 0.1 * m_own).float().cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), (0.9 * rv0.double() + 0.1 * v_own * n / (n - 1)).float().cpu().numpy(), rtol=1e-5, atol=1e-6)
    # (4) normalise + ReLU, tightly: fp32 fma of the kernel's own y_raw and constants, rounded to bf16 once
    own = torch.relu(y_raw.float() * saved[2].view(1, -1, 1, 1) + saved[3].view(1, -1, 1, 1))
    err = (out.float() - own).abs()
    assert bool((err <= 2.0 ** -8 * own.abs() + 1e-6).all()), float(err.max())
    # (5) the block against torch's fp64 training-mode block.  Error budget per element: y_raw's rounding (2^-9 |y| scale), the statistics of the
    # rounded output (1e-3 relative on scale, 1e-3 sigma on the mean, see (2)), the output's own rounding (2^-9 |out|)
    want = torch.relu(F.batch_norm(y64, None, None, gamma.double(), beta.double(), True, 0.0, EPS))
    sc = scale.view(1, -1, 1, 1).abs()
    tol = 2.0 ** -8 * want.abs() + 2.0 ** -8 * y64.abs() * sc + 2e-3 * gamma.double().abs().view(1, -1, 1, 1) + 1e-6
    err = (out.double() - want).abs()
    assert bool((err <= tol).all()), float((err - tol).max())


def test_training_forward_is_deterministic_and_has_no_running_stats_requirement(dev):
    x, w, gamma, beta = _problem(7, 32, 32, seed=5, dev=dev)
    a = _raw_calls(x, w, gamma, beta)
    b = _raw_calls(x, w, gamma, beta)
    torch.cuda.synchronize()
    for s, t in zip(a, b):
        assert torch.equal(s, t)


@pytest.mark.parametrize("case", CASES)
def test_training_backward_stage_by_stage(case, dev):
    B, Hin, Win = case
    x, w, gamma, beta = _problem(B, Hin, Win, seed=B * 7 + Hin, dev=dev)
    g = torch.Generator().manual_seed(B)
    dout = torch.randn(B, 256, 2 * Hin, 2 * Win, generator=g).to(torch.bfloat16).to(dev)
    y_raw, out, saved, dy, dgamma, dbeta, dx = _raw_calls(x, w, gamma, beta, dout=dout)
    torch.cuda.synchronize()
    mean, rstd, scale, shift = (saved[i].double().view(1, -1, 1, 1) for i in range(4))
    # BatchNorm + ReLU backward in fp64 on the kernel's own y_raw / constants
    n = y_raw.numel() // 256
    mask = (y_raw.double() * scale + shift) > 0        # the sign of the kernel's single-rounding fp32 fma is the sign of the exact value
    dz = dout.double() * mask
    xhat = (y_raw.double() - mean) * rstd
    db64 = dz.sum(dim=(0, 2, 3))
    dg64 = (dz * xhat).sum(dim=(0, 2, 3))
    np.testing.assert_allclose(dbeta.cpu().numpy(), db64.cpu().numpy(), rtol=1e-4, atol=1e-4 * float(db64.abs().max()))
    np.testing.assert_allclose(dgamma.cpu().numpy(), dg64.cpu().numpy(), rtol=1e-4, atol=1e-4 * float(dg64.abs().max()))
    dy64 = scale * (dz - db64.view(1, -1, 1, 1) / n - xhat * dg64.view(1, -1, 1, 1) / n)
    err = (dy.double() - dy64).abs()
    assert bool((err <= 2.0 ** -8 * dy64.abs() + 1e-4 * dy64.abs().max()).all()), float(err.max())
    # the input-gradient GEMM, tightly: fp64 stride-2 convolution of the kernel's own dy (4096 products, one bf16 rounding)
    dx64 = F.conv2d(dy.double(), w.double(), stride=2, padding=1)
    err = (dx.double() - dx64).abs()
    assert bool((err <= 2.0 ** -8 * dx64.abs() + 2e-3 * dx64.abs().max()).all()), float(err.max())


def test_input_gradient_impulses_place_every_tap_and_do_not_cross_samples(dev):
    """dy = a single 1 at (b, co, oy, ox): dx[b, ci, iy, ix] must be w[ci, co, oy + 1 - 2 iy, ox + 1 - 2 ix] EXACTLY (one product per output),
    zero elsewhere -- borders, both phase parities, the first row of sample 1 (no leak into the last row of sample 0)."""
    from ihpr_b200._lib import lib, check
    L = lib()
    B, H, W = 2, 32, 32
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(256, 256, 4, 4, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    n = L.ihpr_deconv_train_workspace_bytes(256, 256)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    # an identity BatchNorm state: scale 1, shift +1 with y_raw = 0 -> mask all ones, P = 0 needs dgamma = 0 ... use the GEMM through its own
    # backward entry with saved = (mean 0, rstd 1, scale 1, shift 1): dy_raw = dz - mean(dz) - xhat * mean(dz * xhat) with y_raw = 0 -> xhat = 0,
    # so dy_raw = dz - mean(dz); the test feeds TWO opposite impulses in the same channel so that mean(dz) = 0 and dy_raw = dz exactly.
    saved = torch.zeros((4, 256), dtype=torch.float32, device=dev)
    saved[1:4] = 1.0
    y_raw = torch.zeros((B, 256, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
    stream = torch.cuda.current_stream(dev).cuda_stream
    spots = [(0, 5, 0, 0), (0, 17, 63, 63), (1, 200, 0, 31), (1, 3, 1, 0), (0, 255, 62, 1), (1, 128, 33, 34), (0, 64, 63, 0)]
    for (b, co, oy, ox) in spots:
        dout = torch.zeros((B, 256, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
        dout[b, co, oy, ox] = 1.0
        ob, ooy, oox = 1 - b, (oy + 37) % 64, (ox + 11) % 64
        dout[ob, co, ooy, oox] = -1.0           # cancels the impulse in the channel mean; lands in the OTHER sample
        dy = torch.empty_like(y_raw)
        dgb = torch.empty((2, 256), dtype=torch.float32, device=dev)
        dx = torch.empty((B, 256, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        check(L.ihpr_deconv_bn_relu_train_bwd(dout.data_ptr(), y_raw.data_ptr(), w.data_ptr(), saved.data_ptr(), B, 256, 256, H, W, dy.data_ptr(),
                                              dgb[0].data_ptr(), dgb[1].data_ptr(), dx.data_ptr(), ws.data_ptr(), n, stream))
        torch.cuda.synchronize()
        assert torch.equal(dy, dout)
        # one product per non-zero output, so the result is exact: dx[b, ci, iy, ix] = +-w[ci, co, ky, kx] with oy = 2 iy + ky - 1, ox = 2 ix + kx - 1
        want = torch.zeros((B, 256, H, W), dtype=torch.float32, device=dev)
        for (sb, soy, sox, sign) in ((b, oy, ox, 1.0), (ob, ooy, oox, -1.0)):
            for ky in range(4):
                for kx in range(4):
                    iy2, ix2 = soy + 1 - ky, sox + 1 - kx
                    if iy2 % 2 == 0 and ix2 % 2 == 0 and 0 <= iy2 // 2 < H and 0 <= ix2 // 2 < W:
                        want[sb, :, iy2 // 2, ix2 // 2] += sign * w[:, co, ky, kx].float()
        assert torch.equal(dx.float(), want), (b, co, oy, ox)
        assert int((dx[b] != 0).sum()) > 0


def _wgrad(x, dy):
    """ihpr_deconv_wgrad through the C-ABI: x (B, 256, H, W), dy (B, 256, 2H, 2W) bf16 -> (256, 256, 4, 4) fp32"""
    from ihpr_b200._lib import lib, check
    L = lib()
    dev = x.device
    B, _, H, W = x.shape
    xb = x.contiguous(memory_format=torch.channels_last)
    dyb = dy.contiguous(memory_format=torch.channels_last)
    dw = torch.empty((256, 256, 4, 4), dtype=torch.float32, device=dev)
    n = L.ihpr_deconv_wgrad_workspace_bytes(256, 256)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    check(L.ihpr_deconv_wgrad(xb.data_ptr(), dyb.data_ptr(), B, 256, 256, H, W, dw.data_ptr(), ws.data_ptr(), n, torch.cuda.current_stream(dev).cuda_stream))
    assert L.ihpr_last_launch_count() == 2
    return dw


@pytest.fixture(autouse=True)
def _reset_variant():
    yield
    import ihpr_b200
    ihpr_b200.set_variant(0)


# 0: the default; 21 / 22 / 24: clusters of 1 / 2 / 4 CTAs (TMA multicast of the gradient tiles), 64-pixel stages; 31 / 32 / 34: 32-pixel stages
@pytest.mark.parametrize("variant", [0, 21, 22, 24, 31, 32, 34])
@pytest.mark.parametrize("case", CASES + [(1, 8, 32)])      # (1, 8, 32): 4 tiles per (phase, tap) -- fewer tiles than SMs / 16, the split shrinks
def test_weight_gradient_vs_fp64(case, variant, dev):
    """K11: dW of the transposed convolution against torch's fp64 autograd on the same bf16 operands.  The kernel accumulates bf16 products in
    fp32 (one 64-pixel tile after the other, then the batch splits in index order) and never rounds to bf16: 1e-4 of the largest entry."""
    B, Hin, Win = case
    g = torch.Generator().manual_seed(B * 31 + Hin)
    x = torch.randn(B, 256, Hin, Win, generator=g).to(torch.bfloat16).to(dev)
    dy = torch.randn(B, 256, 2 * Hin, 2 * Win, generator=g).to(torch.bfloat16).to(dev)
    import ihpr_b200
    ihpr_b200.set_variant(variant)
    dw = _wgrad(x, dy)
    dw2 = _wgrad(x, dy)
    torch.cuda.synchronize()
    assert torch.equal(dw, dw2)                  # fixed reduction order
    w0 = torch.zeros(256, 256, 4, 4, dtype=torch.float64, device=dev, requires_grad=True)
    (F.conv_transpose2d(x.double(), w0, stride=2, padding=1) * dy.double()).sum().backward()
    err = float((dw.double() - w0.grad).abs().max() / w0.grad.abs().max())
    assert err <= 1e-4, err


def test_weight_gradient_impulses_place_every_tap(dev):
    """x = a single 1 at (b, ci, iy, ix), dy = a single 1 at (b, co, oy, ox): dW[ci, co, ky, kx] = 1 exactly where oy = 2 iy + ky - 1 and
    ox = 2 ix + kx - 1, 0 everywhere else -- every tap, both parities, the borders, and pairs in different samples (which must give 0)."""
    B, H, W = 2, 32, 32
    spots = [((0, 7, 0, 0), (0, 9, 0, 0)), ((0, 7, 0, 0), (0, 9, 1, 2)), ((1, 100, 31, 31), (1, 3, 63, 63)), ((1, 100, 31, 31), (1, 3, 61, 62)),
             ((0, 255, 16, 5), (0, 0, 31, 9)), ((0, 255, 16, 5), (0, 0, 34, 12)), ((1, 64, 0, 31), (1, 200, 2, 63)),
             ((0, 1, 31, 0), (1, 2, 62, 0)),          # different samples: no contribution at all
             ((0, 1, 5, 5), (0, 2, 20, 20))]          # too far apart: no tap connects them
    for (b, ci, iy, ix), (b2, co, oy, ox) in spots:
        x = torch.zeros((B, 256, H, W), dtype=torch.bfloat16, device=dev)
        dy = torch.zeros((B, 256, 2 * H, 2 * W), dtype=torch.bfloat16, device=dev)
        x[b, ci, iy, ix] = 1.0
        dy[b2, co, oy, ox] = 1.0
        dw = _wgrad(x, dy)
        want = torch.zeros_like(dw)
        ky, kx = oy + 1 - 2 * iy, ox + 1 - 2 * ix
        if b == b2 and 0 <= ky < 4 and 0 <= kx < 4:
            want[ci, co, ky, kx] = 1.0
        torch.cuda.synchronize()
        assert torch.equal(dw, want), ((b, ci, iy, ix), (b2, co, oy, ox), dw.nonzero().tolist()[:8])


def test_autograd_block_vs_torch_fp64_autograd(dev):
    """deconv_bn_relu_train as a differentiable function against torch's fp64 autograd of the same block (bf16-rounded operands).
    Element-wise bounds do not survive ReLU-mask flips of values that round across zero, so gradients are compared in the L2 norm."""
    import ihpr_b200
    B, H, W = 4, 32, 32
    x, w, gamma, beta = _problem(B, H, W, seed=11, dev=dev)
    g = torch.Generator().manual_seed(9)
    dout = torch.randn(B, 256, 2 * H, 2 * W, generator=g).to(torch.bfloat16).to(dev)
    xs = x.float().requires_grad_(True)
    ws_ = w.float().requires_grad_(True)
    gs, bs = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(256, device=dev), torch.ones(256, device=dev)
    out = ihpr_b200.deconv_bn_relu_train(xs, ws_, gs, bs, rm, rv, momentum=0.1, eps=EPS)
    assert out.dtype == torch.bfloat16 and out.shape == (B, 256, 2 * H, 2 * W) and out.is_contiguous(memory_format=torch.channels_last)
    out.backward(dout)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    want = torch.relu(F.batch_norm(F.conv_transpose2d(x64, w64, stride=2, padding=1), None, None, g64, b64, True, 0.0, EPS))
    want.backward(dout.double())

    def rel(a, b):
        return float((a.detach().double() - b.detach()).norm() / b.detach().norm())

    assert rel(out, want) <= 2.0 ** -7
    assert rel(xs.grad, x64.grad) <= 2e-2, rel(xs.grad, x64.grad)
    assert rel(ws_.grad, w64.grad) <= 2e-2, rel(ws_.grad, w64.grad)
    # dgamma / dbeta are sums of ~16 k masked N(0, 1) values per channel (|sum| ~ 90): the ~1e-3 of the elements whose pre-activation lies within
    # the bf16 rounding of y_raw of zero flip their mask, each moving the sum by ~1 -- 1.5e-2 of the norm (measured), the same for ANY bf16 block
    assert rel(gs.grad, g64.grad) <= 3e-2, rel(gs.grad, g64.grad)
    assert rel(bs.grad, b64.grad) <= 3e-2, rel(bs.grad, b64.grad)
    assert xs.grad.dtype == torch.float32 and ws_.grad.shape == w.shape
    assert float(rm.abs().max()) > 0 and rm._version > 0           # updated in place, version bumped for the inference cache


def test_training_block_into_fused_head_vs_oracle_soft_argmax(dev):
    """VERDICT r1 #6's check: the training-mode block feeding K3 (final_layer + soft-argmax on tcgen05) against torch's fp64
    ConvTranspose2d + BatchNorm (batch statistics) + ReLU -> fp64 1x1 conv -> the oracle's soft-argmax (oracle/truth64.c)."""
    import ihpr_b200
    from oracle import truth
    B, J, D = 2, 18, 64
    x, w, gamma, beta = _problem(B, 32, 32, seed=21, dev=dev)
    g = torch.Generator().manual_seed(4)
    wt = (torch.randn(J * D, 256, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = (torch.randn(J * D, generator=g) * 0.5).to(dev)
    with torch.no_grad():
        feat = ihpr_b200.deconv_bn_relu_train(x, w, gamma, beta, None, None, momentum=0.1, eps=EPS)
        coords = ihpr_b200.fused_head_soft_argmax(feat, wt, bias, J)
        f64 = torch.relu(F.batch_norm(F.conv_transpose2d(x.double(), w.double(), stride=2, padding=1), None, None, gamma.double(), beta.double(), True, 0.0, EPS))
        heat = F.conv2d(f64, wt.double().view(J * D, 256, 1, 1), bias.double())
    c64 = truth.soft_argmax_f64(heat.cpu().numpy(), J)[0]
    # the features carry two bf16 roundings (2^-8 relative): the heat-map moves by ~1e-2, the expectation over 64^3 voxels by a few 1e-3 voxel
    assert float(np.abs(coords.cpu().numpy() - c64).max()) <= 2e-2


def test_training_entries_reject_what_they_cannot_do(dev):
    from ihpr_b200._lib import lib
    L = lib()
    x, w, gamma, beta = _problem(1, 8, 32, seed=1, dev=dev)
    n = L.ihpr_deconv_train_workspace_bytes(256, 256)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    y = torch.empty((1, 256, 16, 64), dtype=torch.bfloat16, device=dev)
    saved = torch.empty((4, 256), device=dev)
    args = lambda Cin=256, H=8, W=32, nb=n: (x.data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None, 0.1, EPS, 1, Cin, 256, H, W,  # noqa: E731
                                             y.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), nb, None)
    assert L.ihpr_deconv_bn_relu_train_fwd(*args(Cin=128)) != 0 and b"C_in == C_out == 256" in L.ihpr_last_error()
    assert L.ihpr_deconv_bn_relu_train_fwd(*args(W=24)) != 0
    assert L.ihpr_deconv_bn_relu_train_fwd(*args(H=4)) != 0
    assert L.ihpr_deconv_bn_relu_train_fwd(*args(nb=n - 1)) != 0 and b"workspace" in L.ihpr_last_error()


def test_fused_training_step_matches_the_stock_deconv_stack(dev):
    """ResPoseNet(fused_head=True) in training: deconv blocks 2 and 3 through K9 / K10 (fused_deconv) against the stock cuDNN + BatchNorm stack
    feeding the same fused head -- same loss to bf16 accuracy, parameter gradients close in the L2 norm, running statistics updated alike."""
    import copy
    import types
    import ihpr_b200.model as M
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=64)
    torch.manual_seed(0)
    net_a = M.get_pose_net(cfg, True, 18, fused_head=True, fused_deconv=True).to(dev)
    # a head that is not degenerate: the reference's init (std 0.001) makes every gradient tiny
    with torch.no_grad():
        for m in net_a.head.modules():
            if isinstance(m, (torch.nn.ConvTranspose2d, torch.nn.Conv2d)):
                m.weight.normal_(0, 0.03)
    net_b = copy.deepcopy(net_a)
    net_b.fused_deconv = False
    net_a.train()
    net_b.train()
    B = 4
    img = torch.randn(B, 3, 256, 256, device=dev)
    target = {"coord": torch.rand(B, 18, 3, device=dev) * 63, "vis": (torch.rand(B, 18, device=dev) > 0.2).float(), "have_depth": torch.ones(B, 1, device=dev)}
    losses = []
    for net in (net_a, net_b):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = net(img, target)
        loss.backward()
        losses.append(float(loss))
    assert abs(losses[0] - losses[1]) <= 2e-2 * abs(losses[1]) + 1e-3, losses
    pa, pb = dict(net_a.named_parameters()), dict(net_b.named_parameters())
    for name in ("head.deconv_layers.6.weight", "head.deconv_layers.7.weight", "head.deconv_layers.7.bias", "head.deconv_layers.3.weight",
                 "head.deconv_layers.4.weight", "head.deconv_layers.0.weight", "head.final_layer.weight"):
        ga, gb = pa[name].grad.double(), pb[name].grad.double()
        assert float((ga - gb).norm()) <= 8e-2 * float(gb.norm()) + 1e-12, (name, float((ga - gb).norm() / gb.norm()))
    ba, bb = dict(net_a.named_buffers()), dict(net_b.named_buffers())
    for name in ("head.deconv_layers.7.running_mean", "head.deconv_layers.7.running_var", "head.deconv_layers.4.running_var"):
        np.testing.assert_allclose(ba[name].cpu().numpy(), bb[name].cpu().numpy(), rtol=2e-2, atol=2e-3)
    assert int(ba["head.deconv_layers.7.num_batches_tracked"]) == 1


def test_frozen_batchnorm_in_a_training_head_keeps_the_stock_modules(dev):
    """A BatchNorm2d switched to eval() inside a training head (frozen statistics, what the two-rank equality test does) must normalise with
    its RUNNING statistics and still propagate gradients: HeadNet.features leaves such a block to the stock modules."""
    import types
    import ihpr_b200.model as M
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=8)
    torch.manual_seed(0)
    net = M.get_pose_net(cfg, True, 3, fused_head=True).to(dev)
    net.train()
    for m in net.head.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
            m.running_mean.normal_(0, 0.1)
    x = torch.randn(2, 512, 8, 8, device=dev)
    a = net.head.features(x, fused_training=True)
    b = net.head.deconv_layers(x)
    assert torch.equal(a, b) and a.requires_grad


def _golden_names():
    from conftest import deconv_golden_names
    return deconv_golden_names()


@pytest.mark.parametrize("name", _golden_names())
def test_training_block_vs_reference_golden(name, dev):
    """The training block (K9 kTrain + K10 + K9 kDgrad + K11 through deconv_bn_relu_train) against the fixture written from the REFERENCE'S OWN
    HeadNet block (main/model.py:22-38 in training mode, fp64; oracle/make_golden.py --deconv; the numpy oracle oracle/deconv_block_ref.py is pinned
    to the same fixture by tests/test_oracle.py).  Inputs are regenerated from the fixture's seed (bf16-representable).  Bounds as in BASELINE.md 5:
    two bf16 roundings on the output, statistics of the rounded convolution output, gradient tensors in the L2 norm."""
    import ihpr_b200
    from conftest import load_deconv_golden
    from oracle import deconv_block_ref as R
    g = load_deconv_golden(name)
    B, Cin, Cout, H, W = (int(v) for v in g["shape"])
    x, w, gamma, beta, rm, rv, dout = R.problem(int(g["seed"]), B, Cin, Cout, H, W)
    t = lambda a: torch.from_numpy(a).float().to(dev)      # noqa: E731
    xs, ws = t(x).requires_grad_(True), t(w).requires_grad_(True)
    gs, bs = t(gamma).requires_grad_(True), t(beta).requires_grad_(True)
    rmt, rvt = t(rm), t(rv)
    out = ihpr_b200.deconv_bn_relu_train(xs, ws, gs, bs, rmt, rvt, momentum=0.1, eps=EPS)
    out.backward(t(dout).to(torch.bfloat16))
    torch.cuda.synchronize()
    sigma = np.sqrt(g["var"])
    # running statistics: the batch mean / unbiased variance of the bf16-rounded convolution output enter with momentum 0.1
    assert np.abs(rmt.cpu().numpy() - g["running_mean"]).max() <= 0.1 * 1e-3 * sigma.max() + 1e-6
    np.testing.assert_allclose(rvt.cpu().numpy(), g["running_var"], rtol=0.1 * 2e-3 + 1e-6)
    o = out.detach().double().cpu().numpy()
    want = g["out_sub"]
    # out = relu(xhat * gamma + beta): the rounding of the raw output moves xhat by 2^-9 |y| / sigma <= 2^-8 (|xhat| + |mean| / sigma), the statistics by
    # 1e-3; the result is rounded to bf16 once more
    gam, bet = gamma[None, ::8, None, None], beta[None, ::8, None, None]
    xhat_abs = np.where(want > 0, np.abs(want - bet) / gam, 0.0)
    mean_over_sigma = (np.abs(g["mean"]) / sigma)[None, ::8, None, None]
    err = np.abs(o[:, ::8, ::4, ::4] - want)
    tol = 2.0 ** -8 * np.abs(want) + gam * (2.0 ** -8 * (xhat_abs + mean_over_sigma) + 2e-3 * (1.0 + xhat_abs)) + 1e-6
    assert bool((err <= tol).all()), float((err - tol).max())
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))      # noqa: E731
    assert rel(o.sum(axis=(0, 2, 3)), g["out_sum"]) <= 2e-3
    # gradients in the L2 norm: ~1e-4 of the pre-activations lie within the bf16 rounding of zero and flip their ReLU mask, which alone is
    # sqrt(2e-4) = 1.4e-2 of the norm of every gradient tensor behind the mask; per-channel sums over only 2 k elements scatter more
    assert rel(xs.grad.double().cpu().numpy()[:, ::8, ::2, ::2], g["dx_sub"]) <= 3e-2
    assert rel(xs.grad.double().cpu().numpy().sum(axis=(0, 2, 3)), g["dx_sum"]) <= 3e-2
    assert rel(ws.grad.double().cpu().numpy()[::16, ::16], g["dw_sub"]) <= 3e-2
    assert rel(ws.grad.double().cpu().numpy().sum(axis=(0, 1)), g["dw_tapsum"]) <= 3e-2
    assert rel(gs.grad.double().cpu().numpy(), g["dgamma"]) <= 5e-2
    assert rel(bs.grad.double().cpu().numpy(), g["dbeta"]) <= 5e-2
    # ... and the inference kernel (K9, running statistics) on the UPDATED buffers against the reference block in eval mode
    with torch.no_grad():
        oe = ihpr_b200.deconv_bn_relu(xs.detach(), ws.detach(), gs.detach(), bs.detach(), t(g["running_mean"]), t(g["running_var"]), EPS)
    we = g["out_eval_sub"]
    err = np.abs(oe.double().cpu().numpy()[:, ::8, ::4, ::4] - we)
    assert bool((err <= 2.0 ** -8 * np.abs(we) + 2e-3 * np.abs(we).max()).all()), float(err.max())
