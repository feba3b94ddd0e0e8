"""CPU: host-side mirror of the reference interface -- names, argument checks, error behaviour."""
import sys

import numpy as np
import pytest
import torch


def test_cpu_tensors_are_refused_not_silently_computed(built_lib):
    import ihpr_b200
    with pytest.raises(ihpr_b200.IhprError, match="no CPU fallback"):
        ihpr_b200.soft_argmax(torch.zeros(1, 8, 2, 2), 2)
    with pytest.raises(ihpr_b200.IhprError, match="no CPU fallback"):
        ihpr_b200.JointLocationLoss()(torch.zeros(1, 8, 2, 2), torch.zeros(1, 2, 3), torch.ones(1, 2, 1), torch.ones(1, 1))


def test_reference_assertions_are_kept(built_lib):
    import ihpr_b200
    with pytest.raises(AssertionError):
        ihpr_b200.nets.loss.soft_argmax([1, 2, 3], 2)                  # loss.py:14
    gt = torch.zeros(1, 2, 3, requires_grad=True)
    with pytest.raises(AssertionError, match="gradient w.r.t. targets"):      # loss.py:8-11
        ihpr_b200.JointLocationLoss()(torch.zeros(1, 8, 2, 2), gt, torch.ones(1, 2, 1), torch.ones(1, 1))


def test_shape_errors(built_lib):
    from ihpr_b200 import functional as F
    with pytest.raises(ValueError):
        F._shape(torch.zeros(1, 9, 2, 2), 2)
    with pytest.raises(ValueError):
        F._shape(torch.zeros(9, 2, 2), 3)
    assert F._shape(torch.zeros(2, 12, 5, 7), 3) == (2, 4, 5, 7)


def test_dropin_registers_reference_module_names(built_lib):
    import ihpr_b200
    saved = {k: sys.modules.get(k) for k in ("nets", "nets.loss")}
    try:
        mod = ihpr_b200.install_dropin()
        from nets.loss import soft_argmax, JointLocationLoss, JointMSELoss     # main/train.py:7, common/base.py:71,207
        assert soft_argmax is mod.soft_argmax and JointLocationLoss is ihpr_b200.JointLocationLoss
        assert JointMSELoss is mod.JointMSELoss
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_joint_mse_loss_matches_formula(built_lib):
    import ihpr_b200
    torch.manual_seed(0)
    out, tgt, w = torch.randn(4, 6), torch.randn(4, 2, 3), torch.rand(4, 2, 1)
    got = ihpr_b200.JointMSELoss()(out, tgt, w)
    o, t = out.reshape(4, 2, 3), tgt.reshape(4, 2, 3)
    want = sum(0.5 * ((o[:, j] * w[:, j] - t[:, j] * w[:, j]) ** 2).mean() for j in range(2)) / 2
    assert torch.allclose(got, want, atol=1e-6)


def test_cfg_cross_check(built_lib, monkeypatch):
    import types
    from ihpr_b200.nets import loss
    cfg_mod = types.ModuleType("config")
    cfg_mod.cfg = types.SimpleNamespace(depth_dim=4, output_shape=(2, 2))
    monkeypatch.setitem(sys.modules, "config", cfg_mod)
    with pytest.raises(ValueError, match="disagree with cfg"):
        loss._check_cfg(torch.zeros(1, 10, 2, 2), 2)
    loss._check_cfg(torch.zeros(1, 8, 2, 2), 2)


def test_flip_perm_matches_reference_loop(built_lib):
    """main/test.py:74-75 swaps joints pair by pair with clones; flip_perm must be that sequence as ONE permutation,
    also when pairs share a joint (sequential swaps compose, they do not commute)."""
    import ihpr_b200
    torch.manual_seed(1)
    for pairs in (((1, 4), (2, 5), (3, 6)), ((0, 1), (1, 2)), ((2, 2),), ()):
        f = torch.rand(3, 8, 3)
        ref = f.clone()
        for pair in pairs:
            ref[:, pair[0], :], ref[:, pair[1], :] = ref[:, pair[1], :].clone(), ref[:, pair[0], :].clone()
        assert torch.equal(f[:, ihpr_b200.flip_perm(8, pairs), :], ref)


def test_post_processing_has_no_cpu_path(built_lib):
    import ihpr_b200
    c = torch.rand(2, 4, 3)
    with pytest.raises(ihpr_b200.IhprError):
        ihpr_b200.flip_merge(c, c, 64, ((0, 1),))
    with pytest.raises(ihpr_b200.IhprError):
        ihpr_b200.coords_to_camera(c, bbox=torch.rand(2, 4), center_cam=torch.rand(2, 3), f=torch.rand(2, 2), c=torch.rand(2, 2))


def test_stage_targets_views(built_lib):
    from ihpr_b200.trainer import stage_targets
    coord, vis, hd = torch.rand(4, 5, 3), torch.rand(4, 5, 1), torch.rand(4, 1)
    c2, v2, h2 = stage_targets(coord, vis, hd, torch.device("cpu"))
    assert torch.equal(c2, coord) and torch.equal(v2, vis) and torch.equal(h2, hd)
    assert c2.data_ptr() + coord.numel() * 4 == v2.data_ptr()       # one buffer, back to back


def test_aug_config_and_patch_transform_match_reference_fixtures(built_lib):
    """data.get_aug_config draws what the reference's get_aug_config (dataset.py:184-199) drew under the same seeds, and
    data.gen_trans_from_patch reproduces the transform cv2.getAffineTransform gave the reference (fixtures: oracle/make_golden.py --aug)."""
    import random
    from conftest import aug_golden_names, load_aug_golden
    from ihpr_b200 import data
    from oracle import augment_ref as ar
    for name in aug_golden_names():
        g = load_aug_golden(name)
        for n, seed in enumerate(g["seeds"]):
            np.random.seed(int(seed)); random.seed(int(seed))
            scale, rot, do_flip, color = data.get_aug_config()
            assert (scale, float(rot), do_flip, color) == g["augs"][n]
            bbox, _, _ = ar.synthetic_annotation(g["h"], g["w"], g["J"], int(seed))
            t = data.patch_params(bbox, g["w"], g["augs"][n], tuple(g["input_shape"]))
            assert np.abs(t - g["trans"][n]).max() <= 1e-9
    # inverse map: gen_trans(inv=True) undoes gen_trans
    a = data.gen_trans_from_patch(120.5, 80.25, 200., 150., 256, 256, 1.1, 25., inv=False)
    b = data.gen_trans_from_patch(120.5, 80.25, 200., 150., 256, 256, 1.1, 25., inv=True)
    full = lambda m: np.vstack([m, [0, 0, 1]])          # noqa: E731
    assert np.abs(full(a) @ full(b) - np.eye(3)).max() <= 1e-5


def test_aug_record_layout_is_aligned(built_lib):
    from ihpr_b200 import data
    for B, J in ((1, 1), (3, 17), (32, 18), (5, 2)):
        layout, nbytes = data._record_layout(B, J)
        ends = 0
        for name, (off, dt, n) in layout.items():
            assert off % np.dtype(dt).itemsize == 0 and off >= ends
            ends = off + n * np.dtype(dt).itemsize
        assert ends == nbytes


def test_augment_batch_has_no_cpu_path(built_lib):
    import ihpr_b200
    with pytest.raises(ihpr_b200.IhprError):
        ihpr_b200.augment_batch(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), [[8, 8]], [[0, 0, 8, 8]], np.zeros((1, 2, 3)), np.ones((1, 2)), [ihpr_b200.data.NO_AUG])


def test_deferred_heatmap_answers_shape_questions(built_lib):
    """The stand-in for final_layer's output behaves like the (B, J*D, H, W) tensor for shape queries and
    materialises to exactly the conv it stands for."""
    import ihpr_b200
    from ihpr_b200.nets import loss
    torch.manual_seed(0)
    feat, w, b = torch.randn(2, 8, 4, 6), torch.randn(12, 8, 1, 1), torch.randn(12)
    d = ihpr_b200.DeferredHeatmap(feat, w, b, 3)
    assert d.shape == (2, 12, 4, 6) and d.size() == d.shape and d.size(1) == 12 and d.dim() == 4
    assert d.device == feat.device and d.dtype == feat.dtype and not d.is_cuda and not d.requires_grad
    assert torch.equal(d.materialize(), torch.nn.functional.conv2d(feat, w, b))
    loss._check_cfg(d, 3)                                   # no cfg loaded: passes
    with pytest.raises(ValueError, match="input channels"):
        ihpr_b200.DeferredHeatmap(feat, torch.randn(12, 7, 1, 1), b, 3)
    with pytest.raises(ihpr_b200.IhprError):                # consuming it needs the CUDA path
        loss.soft_argmax(d, 3)
