"""CPU: the model mirror keeps the reference's parameter names / shapes (checkpoint compatibility), and the
one-process-per-GPU trainer logic (batch sharding, mean-of-rank-means, gradient averaging) matches a single
process -- exercised with world_size 2 over gloo.  The CUDA criterion cannot run here, so the trainer is given the
oracle's torch restatement of JointLocationLoss as its criterion (tests may use the oracle; the product never does)."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT


def tiny_cfg():
    return types.SimpleNamespace(resnet_type=18, depth_dim=4, input_shape=(32, 32), output_shape=(8, 8), lr=1e-3,
                                 lr_dec_epoch=[2, 3], lr_dec_factor=0.1, batch_size=4)


def test_state_dict_matches_reference_keys(built_lib):
    from ihpr_b200.model import get_pose_net
    cfg = types.SimpleNamespace(resnet_type=50, depth_dim=64)
    net = get_pose_net(cfg, False, 18)
    ours = {"module." + k: "x".join(str(d) for d in v.shape) for k, v in net.state_dict().items()}
    ref = {}
    for line in open(os.path.join(GOLDEN_DIR, "reference_state_keys_r50_j18.txt")):
        parts = line.rstrip("\n").split(" ")
        ref[parts[0]] = parts[1] if len(parts) > 1 else ""
    assert ours == ref          # names, order-independent, and shapes: the reference's checkpoints load unchanged


def test_forward_contract_and_checkpoint_roundtrip(built_lib, tmp_path):
    from oracle.soft_argmax_ref import RefJointLocationLoss
    from ihpr_b200.model import get_pose_net
    from ihpr_b200.trainer import Trainer, synthetic_batch
    cfg = tiny_cfg()
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, 3)
    tr = Trainer(net, cfg, criterion=RefJointLocationLoss())
    img, coord, vis, hd = synthetic_batch(4, 3, cfg, None, seed=1)
    heat = net(img)                                       # forward(x) -> heat-map, main/model.py:99-103
    assert heat.shape == (4, 3 * cfg.depth_dim, 8, 8)
    l0 = tr.train_step(img, coord, vis, hd)
    assert l0.dim() == 0 and torch.isfinite(l0)
    tr.save(str(tmp_path))
    ck = torch.load(os.path.join(str(tmp_path), "snapshot_0.pth.tar"))
    assert set(ck) == {"epoch", "network", "optimizer", "scheduler"} and all(k.startswith("module.") for k in ck["network"])
    net2 = get_pose_net(cfg, True, 3)
    tr2 = Trainer(net2, cfg, criterion=RefJointLocationLoss())
    tr2.load(os.path.join(str(tmp_path), "snapshot_0.pth.tar"))
    assert tr2.epoch == 1
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)


def _ddp_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.soft_argmax_ref import RefJointLocationLoss
    from ihpr_b200.model import get_pose_net
    from ihpr_b200.trainer import Trainer, global_mean_of_rank_means, shard_range, synthetic_batch
    cfg = tiny_cfg()
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, 3)
    for m in net.modules():                               # batch statistics differ per shard: freeze BN for the comparison
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.0
            m.eval()
    tr = Trainer(net, cfg, criterion=RefJointLocationLoss())
    tr.model.train = lambda *a, **k: tr.model          # keep BN in eval mode inside train_step
    img, coord, vis, hd = synthetic_batch(4, 3, cfg, None, seed=1)
    lo, hi = shard_range(4, rank, world)
    loss = tr.train_step(img[lo:hi], coord[lo:hi], vis[lo:hi], hd[lo:hi])
    g = global_mean_of_rank_means(loss)
    if rank == 0:
        torch.save({"loss": g, "w": tr.raw_model.head.final_layer.weight.detach().clone()}, out)
    dist.destroy_process_group()


def test_two_rank_gloo_step_equals_single_process(built_lib, tmp_path):
    from oracle.soft_argmax_ref import RefJointLocationLoss
    from ihpr_b200.model import get_pose_net
    from ihpr_b200.trainer import Trainer, shard_range, synthetic_batch
    assert shard_range(5, 0, 2) == (0, 2) and shard_range(5, 1, 2) == (2, 5)
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = tiny_cfg()
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, 3)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.0
            m.eval()
    tr = Trainer(net, cfg, criterion=RefJointLocationLoss())
    tr.model.train = lambda *a, **k: tr.model
    img, coord, vis, hd = synthetic_batch(4, 3, cfg, None, seed=1)
    loss = tr.train_step(img, coord, vis, hd)
    # mean of the two rank means == global mean (equal shards, balanced_parallel.py:127); averaged gradients == global gradient
    assert abs(got["loss"].item() - loss.item()) <= 1e-6
    assert torch.allclose(got["w"], net.head.final_layer.weight, atol=1e-6)


def test_epoch_loop_steps_the_schedule_where_the_reference_does_and_resumes(built_lib, tmp_path):
    """main/train.py:45-46 + 91-96 and common/base.py:56-65,109-126: lr * gamma^(milestones <= epoch) during epoch `epoch` (the
    trajectory of the reference's pinned PyTorch 1.0.0), snapshot_{epoch} carries 'epoch': epoch, a resume restarts at epoch + 1 with
    the right learning rate and keeps decaying."""
    from oracle.soft_argmax_ref import RefJointLocationLoss
    from ihpr_b200.model import get_pose_net
    from ihpr_b200.trainer import Trainer, synthetic_batch
    cfg = tiny_cfg()                                   # lr 1e-3, milestones [2, 3], gamma 0.1
    torch.manual_seed(0)
    tr = Trainer(get_pose_net(cfg, True, 3), cfg, criterion=RefJointLocationLoss())
    batch = synthetic_batch(2, 3, cfg, None, seed=1)
    seen = []
    hist = tr.fit(lambda epoch: [batch], end_epoch=5, model_dir=str(tmp_path), log=seen.append)
    lrs = [float(line.split("lr: ")[1].split(" ")[0]) for line in seen]
    assert lrs == pytest.approx([1e-3, 1e-3, 1e-4, 1e-5, 1e-5])
    assert len(hist) == 5 and tr.epoch == 5
    assert sorted(os.listdir(str(tmp_path))) == ["snapshot_%d.pth.tar" % e for e in range(5)]
    assert torch.load(os.path.join(str(tmp_path), "snapshot_3.pth.tar"))["epoch"] == 3
    # resume from the end of epoch 1: epochs 2.. continue on the same trajectory
    tr2 = Trainer(get_pose_net(cfg, True, 3), cfg, criterion=RefJointLocationLoss())
    tr2.load(os.path.join(str(tmp_path), "snapshot_1.pth.tar"))
    assert tr2.epoch == 2 and tr2.current_lr() == pytest.approx(1e-3)
    seen2 = []
    tr2.fit([batch], end_epoch=4, log=seen2.append)
    assert [float(line.split("lr: ")[1].split(" ")[0]) for line in seen2] == pytest.approx([1e-4, 1e-5])
    # a resume under a changed schedule follows the CURRENT config (base.py:115-126)
    cfg3 = tiny_cfg()
    cfg3.lr_dec_epoch = [1, 4]
    tr3 = Trainer(get_pose_net(cfg3, True, 3), cfg3, criterion=RefJointLocationLoss())
    tr3.load(os.path.join(str(tmp_path), "snapshot_1.pth.tar"))
    assert tr3.current_lr() == pytest.approx(1e-4) and tr3.start_epoch() == pytest.approx(1e-4)


def _gather_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ihpr_b200.trainer import gather_coords, shard_range
    full = torch.arange(5 * 3 * 3, dtype=torch.float32).view(5, 3, 3)
    lo, hi = shard_range(5, rank, world)                     # ragged: 2 + 3 samples
    got = gather_coords(full[lo:hi].clone())
    torch.save(got, out + str(rank))
    dist.destroy_process_group()


def test_sharded_inference_gathers_coordinates_not_heatmaps(built_lib, tmp_path):
    """main/test.py:62-65 one process per GPU: every rank keeps its shard's (B_r, J, 3) result and all-gathers it (ragged shards too)."""
    out = str(tmp_path / "g")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gather_worker, args=(2, port, out), nprocs=2, join=True)
    full = torch.arange(5 * 3 * 3, dtype=torch.float32).view(5, 3, 3)
    for r in range(2):
        assert torch.equal(torch.load(out + str(r)), full)


def test_head_features_keeps_the_stock_stack_where_k9_does_not_apply(built_lib):
    """HeadNet.features is deconv_layers(x) bit for bit whenever the tensor-core block (K9: CUDA, eval mode, no autograd, 16- / 32-wide
    map) does not apply -- on the CPU always; and the C-ABI refuses the shapes K9 cannot do without touching a device."""
    from ihpr_b200._lib import lib
    from ihpr_b200.model import get_pose_net
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=4)
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, 3)
    x = torch.randn(2, 512, 8, 8)
    net.eval()
    with torch.no_grad():
        assert torch.equal(net.head.features(x), net.head.deconv_layers(x))
    net.train()
    y = net.head.features(x)
    assert y.requires_grad and y.shape == (2, 256, 64, 64)
    L = lib()
    assert L.ihpr_deconv_bn_relu_workspace_bytes(256, 256) >= 16 * 256 * 256 * 2 + 2 * 256 * 4
    assert L.ihpr_deconv_bn_relu_workspace_bytes(0, 256) == 0
    # argument checks come before any CUDA call: bad channel counts / widths / null pointers are reported, not launched
    assert L.ihpr_deconv_bn_relu(None, None, 1, 256, 256, 32, 32, None, None) < 0
    assert L.ihpr_deconv_bn_relu(256, 256, 1, 256, 128, 32, 32, 256, None) < 0         # C_out must be 256
    assert L.ihpr_deconv_bn_relu(256, 256, 1, 256, 256, 32, 8, 256, None) < 0          # width 8
    assert L.ihpr_deconv_bn_relu(256, 256, 1, 256, 256, 8, 16, 256, None) < 0          # width 16 needs a height that is a multiple of 16


def test_training_deconv_block_falls_back_on_cpu_and_its_entries_check_arguments(built_lib):
    """HeadNet.features(x, fused_training=True) is the stock module stack wherever the training kernels (K9 kTrain / kDgrad, K10, K11: CUDA,
    256 -> 256 channels, 16- / 32-wide map) do not apply -- on the CPU always, with identical running statistics; the training entries of the
    C-ABI report bad arguments before any CUDA call."""
    import copy
    from ihpr_b200._lib import lib
    from ihpr_b200.model import get_pose_net
    cfg = types.SimpleNamespace(resnet_type=18, depth_dim=4)
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, 3, fused_head=True)
    assert net.fused_deconv and not get_pose_net(cfg, True, 3, fused_head=True, fused_deconv=False).fused_deconv
    assert not get_pose_net(cfg, True, 3).fused_deconv                  # only where the fused head is used
    ref = copy.deepcopy(net)
    x = torch.randn(2, 512, 8, 8)
    net.train()
    ref.train()
    y = net.head.features(x, fused_training=True)
    assert torch.equal(y, ref.head.deconv_layers(x)) and y.requires_grad
    for (n, a), (_, b) in zip(net.head.named_buffers(), ref.head.named_buffers()):
        assert torch.equal(a, b), n
    L = lib()
    n = L.ihpr_deconv_train_workspace_bytes(256, 256)
    assert n >= 2 * 16 * 256 * 256 * 2 and L.ihpr_deconv_train_workspace_bytes(0, 256) == 0
    assert L.ihpr_deconv_wgrad_workspace_bytes(256, 256) >= 9 * 16 * 256 * 256 * 4 and L.ihpr_deconv_wgrad_workspace_bytes(128, 256) == 0
    p = 256         # any non-null, aligned value: the checks below fail before a pointer is looked at
    fwd = lambda Cin=256, Cout=256, H=32, W=32, nb=n, rm=None, rv=None: L.ihpr_deconv_bn_relu_train_fwd(  # noqa: E731
        p, p, p, p, rm, rv, 0.1, 1e-5, 1, Cin, Cout, H, W, p, p, p, p, nb, None)
    assert L.ihpr_deconv_bn_relu_train_fwd(None, p, p, p, None, None, 0.1, 1e-5, 1, 256, 256, 32, 32, p, p, p, p, n, None) < 0
    assert fwd(Cin=128) < 0 and b"C_in == C_out == 256" in L.ihpr_last_error()
    assert fwd(Cout=128) < 0 and fwd(W=8) < 0 and fwd(H=12, W=16) < 0
    assert fwd(rm=p) < 0 and b"come together" in L.ihpr_last_error()
    assert fwd(nb=n - 1) < 0 and b"workspace" in L.ihpr_last_error()
    assert L.ihpr_deconv_bn_relu_train_bwd(p, p, None, p, 1, 256, 256, 32, 32, p, p, p, p, p, n, None) < 0        # dx without the weight
    assert L.ihpr_deconv_bn_relu_train_bwd(p, p, p, p, 1, 256, 256, 32, 24, p, p, p, None, p, n, None) < 0
    assert L.ihpr_deconv_wgrad(p, p, 1, 256, 256, 32, 32, None, p, 1 << 30, None) < 0
    assert L.ihpr_deconv_wgrad(p, p, 1, 256, 256, 32, 32, p, p, 1024, None) < 0 and b"workspace" in L.ihpr_last_error()
    assert L.ihpr_deconv_wgrad(p, p, 0, 256, 256, 32, 32, p, p, 1 << 30, None) < 0
