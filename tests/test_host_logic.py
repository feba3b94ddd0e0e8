"""CPU: host-side mirror of the reference interface -- names, argument checks, error behaviour."""
import sys

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
