"""CPU: the oracle restatements against the goldens generated from the reference itself, plus the
analytic known-answer tests of SURVEY.md 8c.  Tolerances are written next to each check."""
import numpy as np
import pytest
import torch

from conftest import deconv_golden_names, golden_names, load_deconv_golden, load_golden, load_post_golden, post_golden_names
from oracle import coords_post_ref, inputs, truth
from oracle.soft_argmax_ref import RefJointLocationLoss, ref_fwd_bwd, ref_soft_argmax


@pytest.mark.parametrize("name", golden_names())
def test_truth64_matches_reference_fp64(name):
    g = load_golden(name)
    loss, coords, grad = truth.fwd_bwd_f64(g["heat"], g["gt"], g["vis"], g["have_depth"])
    assert abs(loss - g["ref64_loss"]) <= 1e-9 * max(1.0, abs(g["ref64_loss"]))
    assert np.abs(coords - g["ref64_coords"]).max() <= 1e-9 * max(g["D"], g["H"], g["W"])
    if "ref64_grad" in g:
        ref = g["ref64_grad"]
        assert np.abs(grad - ref).max() <= 1e-9 * np.abs(ref).max()
    else:
        sub = grad.reshape(-1)[::int(g["grad_stride"])]
        assert np.abs(sub - g["ref64_grad_sub"]).max() <= 1e-9 * g["ref64_grad_max"]
        assert abs(np.abs(grad).sum() - g["ref64_grad_abssum"]) <= 1e-9 * g["ref64_grad_abssum"]


@pytest.mark.parametrize("name", golden_names())
def test_torch_restatement_matches_reference_fp32(name):
    # same ATen op sequence as the reference; bit-identical where the goldens were made, so only
    # summation-order noise (other CPU, other thread count) is allowed: 2e-6 relative.
    g = load_golden(name)
    th, tgt, tvis, thd = (torch.from_numpy(g[k]) for k in ("heat", "gt", "vis", "have_depth"))
    loss, coords, grad = ref_fwd_bwd(th, tgt, tvis, thd)
    assert abs(loss.item() - g["ref32_loss"]) <= 2e-6 * max(1.0, abs(g["ref32_loss"]))
    assert np.abs(coords.numpy() - g["ref32_coords"]).max() <= 2e-6 * max(g["D"], g["H"], g["W"]) * 4
    gn = grad.numpy().reshape(-1)
    ref = g["ref32_grad"].reshape(-1) if "ref32_grad" in g else None
    if ref is None:
        gn = gn[::int(g["grad_stride"])]
        ref = g["ref32_grad_sub"]
    assert np.abs(gn - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-30)


def _coords64(heat, J):
    return truth.soft_argmax_f64(heat, J)[0]


def test_kat_uniform_heatmap_gives_centre():
    B, J, D, H, W = 1, 3, 6, 10, 8
    c = _coords64(np.zeros((B, J * D, H, W), np.float32), J)
    assert np.allclose(c[..., 0], (W - 1) / 2, atol=1e-12)
    assert np.allclose(c[..., 1], (H - 1) / 2, atol=1e-12)
    assert np.allclose(c[..., 2], (D - 1) / 2, atol=1e-12)
    ct = ref_soft_argmax(torch.zeros(B, J * D, H, W), J).numpy()
    assert np.allclose(ct, c, atol=1e-4)


def test_kat_one_hot_peak_and_axis_order():
    B, J, D, H, W = 1, 2, 5, 7, 12      # D != H != W: catches axis mix-ups; order is x(W), y(H), z(D)
    h = np.zeros((B, J, D, H, W), np.float32)
    h[0, 0, 3, 5, 9] = 200.0
    h[0, 1, 1, 6, 2] = 200.0
    c = _coords64(h.reshape(B, J * D, H, W), J)
    assert np.allclose(c[0, 0], [9, 5, 3], atol=1e-9) and np.allclose(c[0, 1], [2, 6, 1], atol=1e-9)
    ct = ref_soft_argmax(torch.from_numpy(h.reshape(B, J * D, H, W)), J).numpy()
    assert np.allclose(ct, c, atol=1e-4)


def test_kat_shift_invariance_and_large_magnitude():
    h = inputs.make_heat("randn3", 1, 2, 8, 8, 8, 0)
    c0 = _coords64(h, 2)
    assert np.abs(_coords64(h + np.float32(512.0), 2) - c0).max() <= 1e-3     # fp32 input rounding only
    big = inputs.make_heat("large", 1, 2, 8, 8, 8, 0)
    assert np.isfinite(_coords64(big, 2)).all()
    assert np.isfinite(ref_soft_argmax(torch.from_numpy(big), 2).numpy()).all()


def test_kat_separable_heatmap():
    D, H, W = 6, 9, 8
    rs = np.random.RandomState(0)
    a, b, c = rs.randn(W), rs.randn(H), rs.randn(D)
    h = (c[:, None, None] + b[None, :, None] + a[None, None, :]).astype(np.float32).reshape(1, D, H, W)
    got = _coords64(h, 1)[0, 0]

    def sa(v):
        p = np.exp(v - v.max()); p /= p.sum()
        return (p * np.arange(len(v))).sum()
    assert np.allclose(got, [sa(a), sa(b), sa(c)], atol=1e-5)


def test_kat_w_flip():
    # main/test.py:73: x' = W - x - 1 for a W-flipped heatmap
    J, D, H, W = 2, 4, 6, 8
    h = inputs.make_heat("randn3", 1, J, D, H, W, 3)
    c = _coords64(h, J)
    cf = _coords64(np.ascontiguousarray(h[..., ::-1]), J)
    assert np.allclose(cf[..., 0], W - 1 - c[..., 0], atol=1e-9)
    assert np.allclose(cf[..., 1:], c[..., 1:], atol=1e-9)


def test_kat_gradient_sums_to_zero_and_masks():
    B, J, D, H, W = 2, 3, 4, 6, 8
    h = inputs.make_heat("randn3", B, J, D, H, W, 4)
    gt, vis, hd = inputs.make_targets(B, J, D, H, W, 4)
    _, _, g = truth.fwd_bwd_f64(h, gt, vis, hd)
    assert np.abs(g.reshape(B * J, -1).sum(1)).max() <= 1e-12
    loss0, _, g0 = truth.fwd_bwd_f64(h, gt, np.zeros_like(vis), hd)
    assert loss0 == 0.0 and np.abs(g0).max() == 0.0
    # have_depth = 0 removes the z term: the loss no longer depends on gt z
    gt2 = gt.copy(); gt2[..., 2] += 5
    l1, _, _ = truth.fwd_bwd_f64(h, gt, vis, np.zeros_like(hd))
    l2, _, _ = truth.fwd_bwd_f64(h, gt2, vis, np.zeros_like(hd))
    assert l1 == l2
    tl = RefJointLocationLoss()(torch.from_numpy(h), torch.from_numpy(gt), torch.from_numpy(vis), torch.zeros(B, 1))
    assert abs(tl.item() - l1) <= 1e-5


def test_fp32_c_port_close_to_truth():
    B, J, D, H, W = 2, 3, 8, 8, 8
    h = inputs.make_heat("randn3", B, J, D, H, W, 5)
    gt, vis, hd = inputs.make_targets(B, J, D, H, W, 5, "rand", "alt")
    loss, coords, g = truth.fwd_bwd_f32_port(h, gt, vis, hd)
    l64, c64, g64 = truth.fwd_bwd_f64(h, gt, vis, hd)
    assert abs(loss - l64) <= 1e-5 and np.abs(coords - c64).max() <= 1e-4
    assert np.abs(g - g64).max() <= 1e-5 * np.abs(g64).max()


def test_targets_must_not_require_grad():
    h = torch.zeros(1, 8, 2, 2)
    gt = torch.zeros(1, 2, 3, requires_grad=True)
    with pytest.raises(AssertionError):
        RefJointLocationLoss()(h, gt, torch.ones(1, 2, 1), torch.ones(1, 1))


@pytest.mark.parametrize("name", post_golden_names())
def test_post_processing_oracle_matches_reference_golden(name):
    """oracle/coords_post_ref.py against fixtures produced by the reference's own warp_coord_to_original / pixel2cam
    (common/utils/pose_utils.py:68-75, 14-20): same numpy statements on the same inputs -> bit-identical."""
    g = load_post_golden(name)
    coords, flipped, bbox, center, f, c = coords_post_ref.make_inputs(g["B"], g["J"], g["D"], g["H"], g["W"], int(g["seed"]), bool(g["flip"]))
    assert np.array_equal(coords, g["coords"]) and np.array_equal(bbox, g["bbox"]) and np.array_equal(f, g["f"])
    merged, pix, cam = coords_post_ref.post_process(coords, flipped, g["pairs"], bbox, center, f, c, g["root"], g["D"], (g["H"], g["W"]),
                                                    float(g["bbox_3d_depth"]))
    assert np.array_equal(merged, g["merged"]) and np.array_equal(pix, g["pixel"]) and np.array_equal(cam, g["cam"])
    if g["root"] >= 0:
        assert np.abs(cam[:, g["root"]]).max() == 0.0


def test_post_processing_known_answers():
    """Centre voxel of the volume -> centre of the box at the box's depth; on the optical axis -> x = y = 0 in camera space."""
    D = H = W = 64
    coords = np.array([[[W / 2, H / 2, D / 2]]], np.float32)
    bbox = np.array([[100., 50., 200., 300.]], np.float32)
    center = np.array([[0., 0., 4000.]], np.float32)
    f = np.array([[1000., 1000.]], np.float32)
    c = np.array([[200., 200.]], np.float32)
    _, pix, cam = coords_post_ref.post_process(coords, None, (), bbox, center, f, c, -1, D, (H, W), 2000.0)
    assert np.allclose(pix[0, 0], [200., 200., 4000.])
    assert np.allclose(cam[0, 0], [0., 0., 4000.])
    # flip merge of a pass with itself mirrored back is the identity
    a = torch.rand(2, 6, 3) * 63
    b = a.clone(); b[:, :, 0] = W - b[:, :, 0] - 1
    assert torch.allclose(coords_post_ref.flip_merge(a, b, W, ()), a, atol=1e-5)


@pytest.mark.parametrize("name", __import__("conftest").aug_golden_names())
def test_augment_oracle_matches_reference_golden(name):
    """oracle/augment_ref.py (numpy restatement incl. OpenCV's fixed-point warpAffine) against what the reference's own
    DatasetLoader.__getitem__ returned: normalised patch bit-identical (sha256 of the fp32 bytes), joints to 1e-6."""
    import hashlib
    from conftest import load_aug_golden
    from oracle import augment_ref as ar
    g = load_aug_golden(name)
    for n, seed in enumerate(g["seeds"]):
        img = ar.synthetic_image(g["h"], g["w"], int(seed))
        assert np.array_equal(np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8), g["src_sha"][n])
        bbox, joints, vis = ar.synthetic_annotation(g["h"], g["w"], g["J"], int(seed))
        o_img, o_joint, o_vis, o_trans = ar.get_item(img, bbox, joints, vis, g["pairs"], g["augs"][n], tuple(g["input_shape"]), tuple(g["output_shape"]),
                                                     g["depth_dim"], float(g["bbox_3d_depth"]), g["pixel_mean"], g["pixel_std"])
        assert np.array_equal(np.frombuffer(hashlib.sha256(o_img.tobytes()).digest(), np.uint8), g["img_sha"][n])
        assert np.array_equal(o_img.reshape(-1)[::97], g["img_sub"][n])
        assert np.abs(o_trans - g["trans"][n]).max() <= 1e-9
        assert np.abs(o_joint - g["joint"][n]).max() <= 1e-4 and np.array_equal(o_vis, g["vis"][n])


@pytest.mark.parametrize("name", deconv_golden_names())
def test_deconv_block_oracle_matches_reference_golden(name):
    """oracle/deconv_block_ref.py (numpy fp64) against the fixture written from the reference's OWN HeadNet block (main/model.py:22-38,
    run in training mode in fp64 by oracle/make_golden.py --deconv): output, running statistics, every gradient."""
    from oracle import deconv_block_ref as R
    g = load_deconv_golden(name)
    B, Cin, Cout, H, W = (int(v) for v in g["shape"])
    x, w, gamma, beta, rm, rv, dout = R.problem(int(g["seed"]), B, Cin, Cout, H, W)
    f = R.forward_train(x, w, gamma, beta, rm, rv)
    b = R.backward_train(x, w, gamma, beta, dout)
    e = R.forward_eval(x, w, gamma, beta, f["running_mean"], f["running_var"])
    pairs = (("running_mean", f["running_mean"]), ("running_var", f["running_var"]), ("dgamma", b["dgamma"]), ("dbeta", b["dbeta"]),
             ("out_sub", f["out"][:, ::8, ::4, ::4]), ("out_eval_sub", e[:, ::8, ::4, ::4]), ("dx_sub", b["dx"][:, ::8, ::2, ::2]),
             ("dw_sub", b["dw"][::16, ::16]), ("out_sum", f["out"].sum(axis=(0, 2, 3))), ("dx_sum", b["dx"].sum(axis=(0, 2, 3))),
             ("dw_tapsum", b["dw"].sum(axis=(0, 1))), ("mean", f["mean"]), ("var", f["var"]))
    for key, val in pairs:
        assert np.abs(val - g[key]).max() <= 1e-9 * max(np.abs(g[key]).max(), 1e-30), key


def test_deconv_block_oracle_known_answers():
    """a single input pixel spreads the 4 x 4 kernel over output rows / columns 2 i - 1 ... 2 i + 2 (torch.nn.ConvTranspose2d, stride 2,
    padding 1), clipped at the border; the gradients of a linear functional are the transposes of that map."""
    from oracle import deconv_block_ref as R
    x = np.zeros((1, 2, 3, 3))
    w = np.arange(2 * 3 * 16, dtype=np.float64).reshape(2, 3, 4, 4)
    x[0, 1, 1, 1] = 1.0
    y = R.conv_transpose2d(x, w)
    assert y.shape == (1, 3, 6, 6)
    assert np.array_equal(y[0, :, 1:5, 1:5], w[1]) and y.sum() == w[1].sum()
    x[:] = 0
    x[0, 0, 0, 0] = 1.0                               # top-left corner: kernel rows / columns 0 fall outside
    y = R.conv_transpose2d(x, w)
    assert np.array_equal(y[0, :, 0:3, 0:3], w[0][:, 1:, 1:]) and y[0, :, 3:, :].sum() == 0
    r = np.random.RandomState(0)
    x, w = r.randn(2, 3, 4, 5), r.randn(3, 2, 4, 4)
    gamma, beta, dout = r.rand(2) + 0.5, r.randn(2), r.randn(2, 2, 8, 10)
    b = R.backward_train(x, w, gamma, beta, dout)
    # finite differences of sum(out * dout) in x and w
    f0 = (R.forward_train(x, w, gamma, beta)["out"] * dout).sum()
    for arr, grad in ((x, b["dx"]), (w, b["dw"])):
        for idx in ((0, 0, 0, 0), (1, 1, 2, 3), (1, 0, 3, 1)):
            old = arr[idx]
            arr[idx] = old + 1e-6
            f1 = (R.forward_train(x, w, gamma, beta)["out"] * dout).sum()
            arr[idx] = old
            assert abs((f1 - f0) / 1e-6 - grad[idx]) <= 1e-4 * max(1.0, abs(grad[idx])), (idx, (f1 - f0) / 1e-6, grad[idx])
