"""GPU: K9 -- ConvTranspose2d(k4 s2 p1, no bias) + BatchNorm2d (eval) + ReLU as one tensor-core kernel (csrc/deconv_bn_relu.cu), the last
deconv block of the reference's HeadNet (main/model.py:22-38) at inference, through the C-ABI.  Truth: torch's own conv_transpose2d /
batch_norm / relu in fp64 on the same bf16-rounded operands.  Bound: the kernel accumulates K = 4 taps x C_in products in fp32 and rounds
the result to bf16 once: |err| <= 2^-8 |y| + 2e-3 max|y|."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _reset_variant():
    yield
    import ihpr_b200
    ihpr_b200.set_variant(0)


def _problem(B, Cin, Hin, seed, Win=32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, Hin, Win, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cin, 256, 4, 4, generator=g) * 0.05).to(torch.bfloat16)
    gamma = torch.rand(256, generator=g) + 0.5
    beta = torch.randn(256, generator=g) * 0.3
    mean = torch.randn(256, generator=g) * 0.2
    var = torch.rand(256, generator=g) + 0.3
    return x, w, gamma, beta, mean, var


def _truth64(x, w, gamma, beta, mean, var, eps, dev):
    y = F.conv_transpose2d(x.to(dev).double(), w.to(dev).double(), stride=2, padding=1)
    y = F.batch_norm(y, mean.to(dev).double(), var.to(dev).double(), gamma.to(dev).double(), beta.to(dev).double(), False, 0.0, eps)
    return torch.relu(y)


@pytest.mark.parametrize("case", [
    (2, 256, 32),       # the head's third block: 32 x 32 -> 64 x 64
    (3, 256, 8),        # one 8-row group per (sample, phase): every tile touches the top AND the bottom border
    (1, 128, 16),       # C_in = 128: two k-blocks per tap
    (37, 256, 32),      # 592 work items: persistent CTAs walk several, the ring wraps across items
    (5, 256, 16, 16),   # the head's second block: 16 x 16 -> 32 x 32 (8 input rows per accumulator, a warp stores 2 rows x 16 pixels)
    (2, 128, 32, 16),   # 16 wide, 32 high: two 16-row groups per (sample, phase)
])
@pytest.mark.parametrize("variant", [0, 21, 22, 24])    # 0: default cluster size; 21 / 22 / 24: clusters of 1 / 2 / 4 CTAs (TMA multicast of the weights)
def test_deconv_bn_relu_vs_torch_fp64(case, variant, dev):
    import ihpr_b200
    B, Cin, Hin = case[:3]
    Win = case[3] if len(case) > 3 else 32
    ihpr_b200.set_variant(variant)
    x, w, gamma, beta, mean, var = _problem(B, Cin, Hin, seed=B * 1000 + Cin + Hin, Win=Win)
    eps = 1e-5
    want = _truth64(x, w, gamma, beta, mean, var, eps, dev)
    with torch.no_grad():
        y = ihpr_b200.deconv_bn_relu(x.to(dev), w.to(dev), gamma.to(dev), beta.to(dev), mean.to(dev), var.to(dev), eps)
    torch.cuda.synchronize()
    assert ihpr_b200.last_launch_count() == 1          # the kernel itself; the parameter preparation was its own call before it
    assert y.shape == (B, 256, 2 * Hin, 2 * Win) and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.double() - want).abs()
    bound = 2.0 ** -8 * want.abs() + 2e-3 * want.abs().max()
    assert bool((err <= bound).all()), (err.max().item(), want.abs().max().item())
    assert (y == 0).float().mean().item() > 0.05        # the ReLU is doing something on this input
    # bit-reproducible, and the NCHW-contiguous input gives the same bits as the channels_last one
    with torch.no_grad():
        y2 = ihpr_b200.deconv_bn_relu(x.to(dev).contiguous(memory_format=torch.channels_last), w.to(dev), gamma.to(dev), beta.to(dev), mean.to(dev),
                                      var.to(dev), eps)
    assert torch.equal(y, y2)
    # the prepared parameters are cached until a parameter tensor changes -- in place counts
    from ihpr_b200 import functional
    wd = w.to(dev)
    args = [t.to(dev) for t in (gamma, beta, mean, var)]
    with torch.no_grad():
        ya = ihpr_b200.deconv_bn_relu(x.to(dev), wd, *args, eps)
        mine = [e for k, e in functional._DECONV_PREPARED.items() if k[2] == id(wd)]
        assert len(mine) == 1
        yb = ihpr_b200.deconv_bn_relu(x.to(dev), wd, *args, eps)
        again = [e for k, e in functional._DECONV_PREPARED.items() if k[2] == id(wd)]
        assert again[0] is mine[0]                      # same tensor objects, unchanged: no second preparation
        args[1].add_(0.25)                              # beta changes in place: the shift must follow
        yc = ihpr_b200.deconv_bn_relu(x.to(dev), wd, *args, eps)
        assert [e for k, e in functional._DECONV_PREPARED.items() if k[2] == id(wd)][0] is not mine[0]
        # a DIFFERENT weight tensor that lands on the recycled address of a freed one must not hit the cache (identity, not address)
        w_other = (wd.float() * 0.5).to(torch.bfloat16)
        addr = w_other.data_ptr()
        y_other = ihpr_b200.deconv_bn_relu(x.to(dev), w_other, *args, eps)
        del w_other
        w_new = wd.clone()                              # very likely the same address
        y_new = ihpr_b200.deconv_bn_relu(x.to(dev), w_new, *args, eps)
        assert torch.equal(y_new, yc) and (w_new.data_ptr() != addr or not torch.equal(y_other, y_new))
    assert torch.equal(ya, y) and torch.equal(yb, y) and not torch.equal(yc, y)
    want_c = torch.relu(F.batch_norm(F.conv_transpose2d(x.to(dev).double(), wd.double(), stride=2, padding=1), mean.to(dev).double(), var.to(dev).double(),
                                     gamma.to(dev).double(), args[1].double(), False, 0.0, eps))
    assert bool(((yc.double() - want_c).abs() <= 2.0 ** -8 * want_c.abs() + 2e-3 * want_c.abs().max()).all())


def test_deconv_impulse_response_places_every_tap(dev):
    """One non-zero input pixel and channel, identity BatchNorm: the output is the 4 x 4 kernel stamped at (2 y - 1, 2 x - 1), clipped at
    the border -- every (phase, tap, shift) of the sub-pixel decomposition is pinned exactly (bf16 values pass through unchanged)."""
    import ihpr_b200
    g = torch.Generator().manual_seed(5)
    w = torch.rand(64, 256, 4, 4, generator=g).to(torch.bfloat16)      # positive: the ReLU is the identity
    one, zero = torch.ones(256), torch.zeros(256)
    for (yy, xx, ci) in ((0, 0, 3), (31, 31, 63), (0, 31, 0), (17, 5, 40), (7, 0, 9), (8, 16, 1)):
        x = torch.zeros(1, 64, 32, 32, dtype=torch.bfloat16)
        x[0, ci, yy, xx] = 1.0
        with torch.no_grad():
            y = ihpr_b200.deconv_bn_relu(x.to(dev), w.to(dev), one.to(dev), zero.to(dev), zero.to(dev), one.to(dev), 0.0)
        want = F.conv_transpose2d(x.float(), w.float(), stride=2, padding=1)
        assert torch.equal(y.float().cpu(), want), (yy, xx, ci)


def test_deconv_rejects_what_the_kernel_cannot_do(dev):
    import ihpr_b200
    x, w, gamma, beta, mean, var = _problem(1, 256, 8, seed=1)
    args = [t.to(dev) for t in (gamma, beta, mean, var)]
    with torch.no_grad():
        with pytest.raises(ihpr_b200.IhprError):
            ihpr_b200.deconv_bn_relu(x.to(dev)[:, :, :, :8], w.to(dev), *args)               # width 8
        with pytest.raises(ihpr_b200.IhprError):
            ihpr_b200.deconv_bn_relu(x.to(dev)[:, :, :4], w.to(dev), *args)                  # height 4
        with pytest.raises(ihpr_b200.IhprError):
            ihpr_b200.deconv_bn_relu(x.to(dev), w.to(dev)[:, :128], *[a[:128] for a in args])    # C_out 128
    with torch.no_grad():
        assert ihpr_b200.deconv_bn_relu(x.to(dev)[:0], w.to(dev), *args).shape == (0, 256, 16, 64)    # empty batch: empty result
    with pytest.raises(ihpr_b200.IhprError):                                                   # forward only
        ihpr_b200.deconv_bn_relu(x.to(dev).requires_grad_(True), w.to(dev), *args)


def test_inference_head_runs_k9_then_k3_and_matches_the_stock_head(dev):
    """ResPoseNet(fused_head=True).predict in eval mode: deconv block 3 through K9, final_layer + soft_argmax through K3 (main/test.py:62-65
    without the (B, 256, 64, 64) fp32 activation or the heat-map) against the stock fp32 module stack + the oracle's soft_argmax."""
    import ihpr_b200
    from ihpr_b200.model import get_pose_net
    from oracle.soft_argmax_ref import ref_soft_argmax
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=32, input_shape=(256, 256), output_shape=(64, 64))
    torch.manual_seed(1)
    J = 5
    net = get_pose_net(cfg, True, J, fused_head=True).to(dev)
    for m in net.head.modules():                       # non-trivial weights and running statistics
        if isinstance(m, torch.nn.ConvTranspose2d):
            torch.nn.init.normal_(m.weight, std=0.03)
        elif isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.normal_(m.bias, std=0.2)
    torch.nn.init.normal_(net.head.final_layer.weight, std=0.02)
    net.eval()
    x = torch.randn(2, 3, 256, 256, device=dev)
    with torch.no_grad():
        feat32 = net.head.deconv_layers(net.backbone(x))
        feat = net.head.features(net.backbone(x))
        assert ihpr_b200.last_launch_count() == 1          # K9 ran (blocks 2 and 3; block 1, 2048 channels on an 8-wide map, is the stock stack)
        assert feat.dtype == torch.bfloat16 and feat.shape == feat32.shape
        rel = (feat.float() - feat32).abs().max().item() / feat32.abs().max().item()
        assert rel <= 2e-2, rel                            # bf16 input and output of the block
        coords = net.predict(x)
        want = ref_soft_argmax(net.head.final_layer(feat32).cpu(), J, cfg.depth_dim)
    assert coords.shape == (2, J, 3)
    assert (coords.cpu() - want).abs().max().item() <= 0.25, (coords.cpu() - want).abs().max().item()
    # training mode / autograd keep the stock stack (batch statistics)
    net.train()
    assert net.head.features(net.backbone(x)).dtype == torch.float32


def test_graphed_predict_replays_the_test_path(dev):
    """GraphedPredict: backbone + head + K9 + K3 (+ the flip pass and K6) captured once, replayed for new batches; same numbers as the
    eager predict() on each batch."""
    from ihpr_b200.model import GraphedPredict, get_pose_net
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=32, input_shape=(256, 256), output_shape=(64, 64))
    torch.manual_seed(3)
    J = 4
    net = get_pose_net(cfg, True, J, fused_head=True).to(dev)
    torch.nn.init.normal_(net.head.final_layer.weight, std=0.02)
    with pytest.raises(ValueError):
        GraphedPredict(net, torch.randn(4, 3, 256, 256, device=dev))      # still in training mode
    net.eval()
    pairs = ((0, 1), (2, 3))
    gp = GraphedPredict(net, torch.randn(4, 3, 256, 256, device=dev), flip_pairs=pairs)
    for seed in (10, 11, 12):
        img = torch.randn(4, 3, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
        got = gp(img).clone()
        with torch.no_grad():
            want = net.predict(img, flip_pairs=pairs)
        assert got.shape == (4, J, 3)
        assert (got - want).abs().max().item() <= 1e-3, (got - want).abs().max().item()
    with pytest.raises(ValueError):
        gp(torch.randn(2, 3, 256, 256, device=dev))
