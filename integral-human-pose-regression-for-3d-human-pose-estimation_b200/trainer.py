"""One-process-per-GPU trainer: the replacement for the reference's single-process DataParallelModel /
DataParallelCriterion (common/nets/balanced_parallel.py:58-183, common/base.py:68-134).

* the batch is sharded by rank (what DataParallel.scatter did per step, base.py:96), parameters live on every rank
  (no per-step broadcast), gradients are averaged with an NCCL all-reduce overlapped with backward (torch DDP);
* the loss of a step is the mean of the per-rank means, which is what `Reduce.apply(*outputs) / len(outputs)`
  computes (balanced_parallel.py:127) and equals the global mean for equal shards;
* Adam(lr) + MultiStepLR(lr_dec_epoch, lr_dec_factor) as in base.py:75-85 / main/config.py:35-39, stepped where the reference
  steps it (top of every epoch, main/train.py:45-46): `fit()` / `start_epoch()` / `end_epoch()`; resume = base.py:56-65,109-126;
* inference shards the batch and gathers (B, J, 3) coordinates instead of heat-maps (`predict_sharded`, main/test.py:62-65);
* checkpoints keep the reference's layout: {'epoch', 'network', 'optimizer', 'scheduler'} with `module.`-prefixed
  network keys (main/train.py:91-96, base.py:51-65).
"""
import contextlib
import os
import types

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

DEFAULT_CFG = types.SimpleNamespace(resnet_type=50, depth_dim=64, input_shape=(256, 256), output_shape=(64, 64),
                                    lr=1e-3, lr_dec_epoch=[210, 280], lr_dec_factor=0.1, batch_size=32)


_null = contextlib.nullcontext


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


def gather_coords(local_coords, group=None):
    """All ranks' (B_r, J, 3) coordinates -> one (sum B_r, J, 3) tensor on every rank, rank order = batch order.  This is what
    replaces the reference's gather of full (B, J*D, H, W) heat-maps onto GPU 0 (main/test.py:62-65 through
    DataParallelModel.gather, balanced_parallel.py:96-99): 12*J bytes per sample on the wire instead of 4*J*D*H*W.
    Shards may be ragged (the last test batch): sizes are exchanged first and shorter shards padded for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_coords
    world = dist.get_world_size(group)
    n = torch.tensor([local_coords.shape[0]], device=local_coords.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(t.item()) for t in sizes]
    cap = max(sizes)
    padded = local_coords.new_zeros((cap,) + tuple(local_coords.shape[1:]))
    padded[:local_coords.shape[0]] = local_coords
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous(), group=group)
    return torch.cat([t[:k] for t, k in zip(out, sizes)], dim=0)


def predict_sharded(model, input_img_local, flip_pairs=None, group=None):
    """Multi-GPU inference of main/test.py:56-76, one process per GPU: every rank runs the network and the soft-argmax (and the
    flip-test merge) on ITS shard of the test batch, then the (B_local, J, 3) results are all-gathered."""
    net = model.module if isinstance(model, DDP) else model
    with torch.no_grad():
        coords = net.predict(input_img_local, flip_pairs=flip_pairs)
    return gather_coords(coords, group)


class Trainer:
    def __init__(self, model, cfg=DEFAULT_CFG, criterion=None, device=None, autocast_dtype=None, channels_last=False, static_graph=True,
                 graph_capture=False):
        """graph_capture=True (multi-GPU only matters): build DDP on the side stream that capture() will later capture on -- PyTorch's
        recipe for CUDA graphs under DDP (its gradient-accumulation nodes remember the stream they were created on)."""
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
        self._side = torch.cuda.Stream(self.device) if (graph_capture and self.device.type == "cuda") else None
        if self.world > 1:
            ids = [self.device.index] if self.device.type == "cuda" else None
            # BN running statistics stay per rank, as in the reference's replicas (no SyncBN, balanced_parallel.py:16-43 is dead code)
            ctx = torch.cuda.stream(self._side) if self._side is not None else _null()
            if self._side is not None:
                self._side.wait_stream(torch.cuda.current_stream(self.device))
            with ctx:
                self.model = DDP(model, device_ids=ids, gradient_as_bucket_view=True, bucket_cap_mb=25, broadcast_buffers=False,
                                 static_graph=static_graph)
            if self._side is not None:
                torch.cuda.current_stream(self.device).wait_stream(self._side)
        else:
            self.model = model
        # base.py:75-77 (Adam, lr from the config); the fused multi-tensor implementation when the parameters are on a GPU.
        # On a GPU the learning rate is a DEVICE TENSOR: MultiStepLR updates it in place (fill_), so a captured CUDA graph of the
        # step (capture()) sees every later decay without being re-captured.
        on_gpu = self.device.type == "cuda"
        lr = torch.tensor(float(cfg.lr), device=self.device) if on_gpu else cfg.lr
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, fused=on_gpu, capturable=on_gpu)
        self.scheduler = torch.optim.lr_scheduler.MultiStepLR(self.optimizer, milestones=list(cfg.lr_dec_epoch),
                                                              gamma=cfg.lr_dec_factor)               # base.py:83-85
        self.autocast_dtype = autocast_dtype
        self.epoch = 0          # the epoch being (or about to be) trained = the reference's loop variable (main/train.py:45)
        self._graph = None

    # ---- epoch loop of main/train.py:45-96 ---------------------------------------------------------------------------
    def current_lr(self):
        lr = self.optimizer.param_groups[0]["lr"]
        return float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr)

    def start_epoch(self):
        """`trainer.scheduler.step()` at the top of every epoch (main/train.py:46).  The reference is pinned to PyTorch 1.0.0
        (README.md:22-24), where MultiStepLR starts at last_epoch = -1 and that call makes last_epoch == epoch, i.e. the epoch
        trains with lr * gamma^(milestones <= epoch).  Current PyTorch already sits at last_epoch = 0 after construction, so the
        same trajectory is: step until last_epoch == epoch (no call for epoch 0, one call per later epoch, and the right
        number of calls after a resume)."""
        while self.scheduler.last_epoch < self.epoch:
            self.scheduler.step()
        return self.current_lr()

    def end_epoch(self, model_dir=None):
        """End of the reference's epoch body: write snapshot_{epoch}.pth.tar with 'epoch': epoch (main/train.py:91-96), then advance."""
        if model_dir is not None:
            self.save(model_dir)
        self.epoch += 1

    def fit(self, batches, end_epoch, model_dir=None, step_fn=None, log=None):
        """main/train.py:45-96: for epoch in range(start_epoch, end_epoch): scheduler.step(); one pass over `batches`; snapshot.
        `batches` is a callable epoch -> iterable of (input_img, joint_img, joint_vis, joints_have_depth) for this rank, or such
        an iterable.  Returns the last loss of every epoch (rank-local, detached)."""
        step_fn = step_fn or self.train_step
        history = []
        while self.epoch < end_epoch:
            lr = self.start_epoch()
            last = None
            for batch in (batches(self.epoch) if callable(batches) else batches):
                last = step_fn(*batch)
            if log is not None:
                log("Epoch %d/%d lr: %g loss_loc: %s" % (self.epoch, end_epoch, lr, "-" if last is None else "%.4f" % float(last)))
            history.append(last)
            self.end_epoch(model_dir)
        return history

    def train_step(self, input_img, joint_img, joint_vis, joints_have_depth):
        """main/train.py:54-72 for this rank's shard.  Returns the (device) loss of this rank."""
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        if self.channels_last:
            input_img = input_img.contiguous(memory_format=torch.channels_last)
        loss = self._forward_backward(self.model, input_img, joint_img, joint_vis, joints_have_depth)
        self.optimizer.step()
        return loss

    def _forward_backward(self, model, input_img, joint_img, joint_vis, joints_have_depth):
        raw = self.raw_model
        step = getattr(getattr(raw, "criterion", None), "forward_backward", None)
        ctx = torch.autocast(device_type=self.device.type, dtype=self.autocast_dtype) if self.autocast_dtype is not None else _null()
        if step is not None and not getattr(raw, "fused_head", False) and self.device.type == "cuda":
            # main/train.py:64-71 with the stored heat-map: model -> criterion + backward in ONE launch (K5), its gradient handed
            # straight to autograd (no ones-fill, no rescale launch)
            with ctx:
                heat = model(input_img)
            return step(heat, joint_img, joint_vis, joints_have_depth)
        target = {"coord": joint_img, "vis": joint_vis, "have_depth": joints_have_depth}
        with ctx:
            loss = model(input_img, target)
        loss.backward()
        return loss.detach()

    # ---- whole-step CUDA graph (single GPU): ~1200 launches per step become one graph replay ------------------------
    def capture(self, input_img, joint_img, joint_vis, joints_have_depth, warmup=3):
        """Capture forward + loss + backward + Adam step into one CUDA graph on static input buffers (PyTorch's
        whole-network capture recipe).  The sm_100a ops are capture-safe: they only enqueue on the current stream,
        take plain device pointers and allocate nothing themselves.  Afterwards `graphed_step(batch)` copies a batch
        into the static buffers and replays.  The learning rate is a device tensor (see __init__), so scheduler steps after the
        capture are seen by the replays.  Under DDP (world > 1) the bucketed NCCL all-reduces are captured with the step: PyTorch's
        recipe asks for >= 11 eager DDP iterations on the side stream first and for the process group to have been created with
        TORCH_NCCL_ASYNC_ERROR_HANDLING=0 (bench.py sets it before init_process_group when --cuda-graph is given)."""
        assert self.device.type == "cuda"
        if self.world > 1:
            warmup = max(warmup, 11)
        self._static = [t.clone() for t in (input_img, joint_img, joint_vis, joints_have_depth)]
        if self.channels_last:
            self._static[0] = self._static[0].contiguous(memory_format=torch.channels_last)
        side = self._side if self._side is not None else torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.train_step(*self._static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        self._graph = torch.cuda.CUDAGraph()

        with torch.cuda.graph(self._graph, stream=side):
            loss = self._forward_backward(self.model, *self._static)
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
        """Resume as common/base.py:56-65 + 109-126 do: network / optimizer / scheduler states, start_epoch = ckpt['epoch'] + 1, then
        the schedule is re-pointed at the CURRENT config (milestones, gamma) and the learning rate re-derived from it
        (lr * gamma^(milestones <= start_epoch); the reference assigns that value to `optimizer.lr`, an attribute Adam never
        reads -- here it reaches the parameter groups).  The next start_epoch() then steps the scheduler like any other epoch."""
        ckpt = torch.load(path, map_location=map_location or self.device)
        load_reference_network(self.raw_model, ckpt["network"])
        lr_obj = self.optimizer.param_groups[0]["lr"]
        if "optimizer" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer"])
        if "scheduler" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler"])
        self.epoch = int(ckpt.get("epoch", -1)) + 1                                                  # base.py:62
        from collections import Counter
        self.scheduler.milestones = Counter(list(self.cfg.lr_dec_epoch))                             # base.py:115-118
        self.scheduler.gamma = self.cfg.lr_dec_factor
        last = self.epoch - 1               # the epoch the snapshot finished = scheduler.last_epoch at save time
        self.scheduler.last_epoch = max(last, 0)
        lr_now = float(self.cfg.lr) * float(self.cfg.lr_dec_factor) ** sum(1 for m in self.cfg.lr_dec_epoch if m <= max(last, 0))
        for gr in self.optimizer.param_groups:
            if isinstance(lr_obj, torch.Tensor):        # keep the SAME device tensor: a captured graph reads it
                lr_obj.fill_(lr_now)
                gr["lr"] = lr_obj
            else:
                gr["lr"] = lr_now
        self.scheduler._last_lr = [self.current_lr() for _ in self.optimizer.param_groups]
        return ckpt


def load_reference_network(model, network_state):
    """`ckpt['network']` of the reference carries DataParallel's `module.` prefix (main/train.py:93)."""
    clean = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in network_state.items()}
    return model.load_state_dict(clean, strict=True)
