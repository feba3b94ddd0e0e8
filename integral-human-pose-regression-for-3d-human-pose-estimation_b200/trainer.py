"""One-process-per-GPU trainer: the replacement for the reference's single-process DataParallelModel /
DataParallelCriterion (common/nets/balanced_parallel.py:58-183, common/base.py:68-134).

* the batch is sharded by rank (what DataParallel.scatter did per step, base.py:96), parameters live on every rank
  (no per-step broadcast), gradients are averaged with an NCCL all-reduce overlapped with backward (torch DDP);
* the loss of a step is the mean of the per-rank means, which is what `Reduce.apply(*outputs) / len(outputs)`
  computes (balanced_parallel.py:127) and equals the global mean for equal shards;
* Adam(lr) + MultiStepLR(lr_dec_epoch, lr_dec_factor) as in base.py:75-85 / main/config.py:35-39;
* checkpoints keep the reference's layout: {'epoch', 'network', 'optimizer', 'scheduler'} with `module.`-prefixed
  network keys (main/train.py:91-96, base.py:51-65).
"""
import os
import types

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

DEFAULT_CFG = types.SimpleNamespace(resnet_type=50, depth_dim=64, input_shape=(256, 256), output_shape=(64, 64),
                                    lr=1e-3, lr_dec_epoch=[210, 280], lr_dec_factor=0.1, batch_size=32)


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of a batch of n samples owned by `rank`."""
    return n * rank // world, n * (rank + 1) // world


def global_mean_of_rank_means(local_mean, group=None):
    """Reduce.apply(...)/N of the reference: average of the per-rank losses (for logging; detached)."""
    t = local_mean.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= dist.get_world_size(group)
    return t


def synthetic_batch(batch, joint_num, cfg, device, seed, pin=False):
    """Shapes and units of data/dataset.py:146-152: image (B,3,H,W), joint_img (B,J,3) in heat-map voxel units,
    joint_vis (B,J,1), joints_have_depth (B,1)."""
    g = torch.Generator().manual_seed(seed)
    H, W = cfg.input_shape
    img = torch.randn(batch, 3, H, W, generator=g)
    coord = torch.rand(batch, joint_num, 3, generator=g) * torch.tensor([cfg.output_shape[1], cfg.output_shape[0], cfg.depth_dim], dtype=torch.float32)
    vis = (torch.rand(batch, joint_num, 1, generator=g) > 0.1).float()
    have_depth = torch.ones(batch, 1)
    out = [img, coord, vis, have_depth]
    if pin:
        out = [t.pin_memory() for t in out]
    return [t.to(device, non_blocking=True) for t in out] if device is not None else out


def stage_targets(joint_img, joint_vis, joints_have_depth, device):
    """One pinned record and ONE host->device copy for the three target tensors of a batch (data/dataset.py:146-152 hands
    them over separately and main/train.py:57-59 copies them one by one).  Returns device views (coord (B,J,3),
    vis (B,J,1), have_depth (B,1)) into the single staged buffer; the C-ABI takes them as plain pointers."""
    B, J = joint_img.shape[0], joint_img.shape[1]
    n = B * J * 3 + B * J + B
    host = torch.empty(n, dtype=torch.float32, pin_memory=(device.type == "cuda"))
    host[:B * J * 3] = joint_img.reshape(-1)
    host[B * J * 3:B * J * 4] = joint_vis.reshape(-1)
    host[B * J * 4:] = joints_have_depth.reshape(-1)
    dev = host.to(device, non_blocking=True)
    return dev[:B * J * 3].view(B, J, 3), dev[B * J * 3:B * J * 4].view(B, J, 1), dev[B * J * 4:].view(B, 1)


class Trainer:
    def __init__(self, model, cfg=DEFAULT_CFG, criterion=None, device=None, autocast_dtype=None, channels_last=False):
        self.cfg = cfg
        self.device = device if device is not None else torch.device("cpu")
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        model = model.to(self.device)
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.channels_last = channels_last
        if criterion is not None:
            model.criterion = criterion
        self.raw_model = model
        if self.world > 1:
            ids = [self.device.index] if self.device.type == "cuda" else None
            # BN running statistics stay per rank, as in the reference's replicas (no SyncBN, balanced_parallel.py:16-43 is dead code)
            self.model = DDP(model, device_ids=ids, gradient_as_bucket_view=True, bucket_cap_mb=25, broadcast_buffers=False,
                             static_graph=True)
        else:
            self.model = model
        # base.py:75-77 (Adam, lr from the config); the fused multi-tensor implementation when the parameters are on a GPU
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=cfg.lr, fused=(self.device.type == "cuda"),
                                          capturable=(self.device.type == "cuda"))
        self.scheduler = torch.optim.lr_scheduler.MultiStepLR(self.optimizer, milestones=list(cfg.lr_dec_epoch),
                                                              gamma=cfg.lr_dec_factor)               # base.py:83-85
        self.autocast_dtype = autocast_dtype
        self.epoch = 0

    def train_step(self, input_img, joint_img, joint_vis, joints_have_depth):
        """main/train.py:54-72 for this rank's shard.  Returns the (device) loss of this rank."""
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        if self.channels_last:
            input_img = input_img.contiguous(memory_format=torch.channels_last)
        target = {"coord": joint_img, "vis": joint_vis, "have_depth": joints_have_depth}
        if self.autocast_dtype is not None:
            with torch.autocast(device_type=self.device.type, dtype=self.autocast_dtype):
                loss = self.model(input_img, target)
        else:
            loss = self.model(input_img, target)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    # ---- whole-step CUDA graph (single GPU): ~1200 launches per step become one graph replay ------------------------
    def capture(self, input_img, joint_img, joint_vis, joints_have_depth, warmup=3):
        """Capture forward + loss + backward + Adam step into one CUDA graph on static input buffers (PyTorch's
        whole-network capture recipe).  The sm_100a ops are capture-safe: they only enqueue on the current stream,
        take plain device pointers and allocate nothing themselves.  Afterwards `graphed_step(batch)` copies a batch
        into the static buffers and replays.  Not combined with DDP here."""
        assert self.world == 1 and self.device.type == "cuda"
        self._static = [t.clone() for t in (input_img, joint_img, joint_vis, joints_have_depth)]
        if self.channels_last:
            self._static[0] = self._static[0].contiguous(memory_format=torch.channels_last)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.train_step(*self._static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        self._graph = torch.cuda.CUDAGraph()
        target = {"coord": self._static[1], "vis": self._static[2], "have_depth": self._static[3]}
        with torch.cuda.graph(self._graph):
            if self.autocast_dtype is not None:
                with torch.autocast(device_type="cuda", dtype=self.autocast_dtype):
                    loss = self.model(self._static[0], target)
            else:
                loss = self.model(self._static[0], target)
            loss.backward()
            self.optimizer.step()
        self._static_loss = loss.detach()
        return self

    def graphed_step(self, input_img, joint_img, joint_vis, joints_have_depth):
        for dst, src in zip(self._static, (input_img, joint_img, joint_vis, joints_have_depth)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_loss

    # ---- checkpoints in the reference's format ------------------------------------------------------------------
    def state(self):
        net = {("module." + k): v for k, v in self.raw_model.state_dict().items()}
        return {"epoch": self.epoch, "network": net, "optimizer": self.optimizer.state_dict(), "scheduler": self.scheduler.state_dict()}

    def save(self, model_dir):
        if self.rank == 0:
            os.makedirs(model_dir, exist_ok=True)
            torch.save(self.state(), os.path.join(model_dir, "snapshot_%d.pth.tar" % self.epoch))     # base.py:51-54

    def load(self, path, map_location=None):
        ckpt = torch.load(path, map_location=map_location or self.device)
        load_reference_network(self.raw_model, ckpt["network"])
        if "optimizer" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer"])
        if "scheduler" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler"])
        self.epoch = int(ckpt.get("epoch", -1)) + 1                                                  # base.py:62
        return ckpt


def load_reference_network(model, network_state):
    """`ckpt['network']` of the reference carries DataParallel's `module.` prefix (main/train.py:93)."""
    clean = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in network_state.items()}
    return model.load_state_dict(clean, strict=True)
