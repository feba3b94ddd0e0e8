#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU, weak scaling

A step = one forward + backward of the integral L1 loss (soft-argmax + JointLocationLoss,
/root/reference/common/nets/loss.py:13-52 and its autograd backward) over one batch of synthetic
heatmaps: B=32 per GPU, J=18, D=H=W=64, fp32 -- the head output of BASELINE.json configs[1].
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "soft-argmax fwd+bwd joint-volumes/s"
UNIT = "joint-volumes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--joints", type=int, default=18)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 10)")
    ap.add_argument("--slices", type=int, default=16)
    ap.add_argument("--workload", default="path", choices=["path", "train"],
                    help="path: the soft-argmax/loss hot path (headline); train: ResNet + head + integral loss training step")
    ap.add_argument("--resnet", type=int, default=50)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--unfused-loss", action="store_true", help="train workload: K1 + K2 instead of K5")
    ap.add_argument("--fused-head", action="store_true", help="train workload: final_layer + loss as K3/K4 (heat-map never stored)")
    ap.add_argument("--stock-deconv", action="store_true", help="train workload with --fused-head: keep deconv blocks 2 / 3 on cuDNN + torch BatchNorm "
                                                                "(comparison arm; default: K9 / K10 training kernels, row N1)")
    ap.add_argument("--cuda-graph", action="store_true", help="train workload, 1 GPU: replay the whole step as one CUDA graph")
    ap.add_argument("--torch-loss", action="store_true", help="train workload: the reference's eager torch loss on the GPU (comparison arm)")
    ap.add_argument("--no-train", action="store_true", help="path workload: skip the `train` sub-record (ResNet-50 training step, BASELINE configs[1-3])")
    ap.add_argument("--train-steps", type=int, default=20)
    return ap.parse_args()


def workload_name(B, J, D, W, dtype):
    return ("integral-L1 soft-argmax fwd+bwd (JointLocationLoss + backward), B=%d per GPU, J=%d, D=%d, H=W=%d, %s heatmaps"
            % (B, J, D, W, dtype))


def make_config(B, J, D, W, dtype):
    """`config`: the SAME dict in both arms (the driver compares the arms on metric and config); arm-specific facts go to `notes`."""
    es = 4 if dtype == "f32" else 2
    mib = (B * J * D * W * W * es) >> 20
    return {"workload": workload_name(B, J, D, W, dtype),
            "l2": "inputs %d MiB + gradients %d MiB per step and GPU >> 126 MB L2: no flush needed between timed iterations" % (mib, mib)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs_torch(B, J, D, H, W, device, dtype, seed):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    heat = torch.randn(B, J * D, H, W, device=device, generator=g, dtype=torch.float32).to(dtype)
    gt = torch.rand(B, J, 3, device=device, generator=g) * torch.tensor([W, H, D], device=device, dtype=torch.float32)
    vis = torch.ones(B, J, 1, device=device)
    hd = torch.ones(B, 1, device=device)
    return heat, gt, vis, hd


def load_cpu_reference():
    """The reference's own CPU implementation of the path: common/nets/loss.py:13-52 UNMODIFIED from baseline/_ref (placed there by
    __graft_entry__.build() / oracle/load_reference.install(); git-ignored, travels with the snapshot), run through the two-attribute
    CPU shim.  Falls back to the oracle port (bit-identical where the goldens were made) only when the copy is absent.
    Returns (step(heat, gt, vis, hd), kind, description)."""
    from oracle import load_reference as lr
    root = lr.installed()
    if root is not None:
        try:
            ref = lr.Reference(root)
            return ref.step, "reference", "the reference's JointLocationLoss.forward + autograd backward (baseline/_ref/common/nets/loss.py, unmodified)"
        except Exception as e:      # noqa: a broken copy must not take the bench down
            sys.stderr.write("bench: baseline/_ref present but not loadable (%r); timing the oracle port instead\n" % (e,))
    from oracle.soft_argmax_ref import ref_fwd_bwd
    return (lambda h, g, v, d: ref_fwd_bwd(h, g, v, d)[0]), "port", "oracle/soft_argmax_ref.py (op-for-op port of loss.py:13-52; baseline/_ref absent)"


def cpu_reference_leg(step, B, J, D, H, W, min_seconds, warmup=2, max_iters=100000):
    """`step` on a bounded sample of the workload, all host threads torch has.  Returns (volumes/s, cores, iters, seconds)."""
    import torch
    cores = torch.get_num_threads()
    heat, gt, vis, hd = make_inputs_torch(B, J, D, H, W, "cpu", torch.float32, 0)
    for _ in range(warmup):
        step(heat, gt, vis, hd)
    n, t0 = 0, time.perf_counter()
    while True:
        step(heat, gt, vis, hd)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or n >= max_iters:
            break
    return B * J * n / dt, cores, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    B, J, D, W = args.batch, args.joints, args.depth, args.hw
    step, kind, what = load_cpu_reference()
    heat, gt, vis, hd = make_inputs_torch(B, J, D, W, W, "cpu", torch.float32, 0)      # the FULL batch of the workload, every step
    for _ in range(max(1, min(args.warmup, 3))):
        step(heat, gt, vis, hd)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step(heat, gt, vis, hd)
        done += 1
        if time.perf_counter() - t0 > 150:      # keep the whole run within minutes on a slow host
            break
    dt = time.perf_counter() - t0
    val = B * J * done / dt
    cores = torch.get_num_threads()
    sample = "the full B=%d batch per step (J=%d, %dx%dx%d fp32), %d steps; %s; host has %s logical CPUs" % (B, J, D, W, W, done, what, os.cpu_count())
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(B, J, D, W, "f32"),
        "notes": {"residency": "host memory (CPU arm, rank 0 only)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def measure_pcie(dev, nbytes=256 << 20, reps=3):
    """Pinned host <-> device copy bandwidth of THIS box (GB/s), one direction at a time: the floor e2e can be held against."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = []
    for src, dst in ((h, d), (d, h)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out.append(nbytes / (best * 1e-3) / 1e9)
    # both directions at once (what the pipelined step does): two streams, each direction timed by its own events
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s1):
            ev[0].record(); d.copy_(h, non_blocking=True); ev[1].record()
        with torch.cuda.stream(s2):
            ev[2].record(); h2.copy_(d2, non_blocking=True); ev[3].record()
        torch.cuda.synchronize()
        best = min(best, max(ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])))
    out.append(nbytes / (best * 1e-3) / 1e9)
    return out[0], out[1], out[2]


def gather_list(x, world):
    import torch.distributed as dist
    if world == 1:
        return [x]
    out = [None] * world
    dist.all_gather_object(out, x)
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import ihpr_b200
    from ihpr_b200 import functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA sm_100 device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if not args.no_train:
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")        # PyTorch's recipe for capturing DDP's NCCL work (train.cuda_graph)
        dist.init_process_group("nccl", device_id=dev)

    B, J, D, W = args.batch, args.joints, args.depth, args.hw
    H = W
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if args.dtype == "f32" else 2
    R, N = B * J, D * H * W
    ihpr_b200.set_variant(args.variant)
    heat, gt, vis, hd = make_inputs_torch(B, J, D, H, W, dev, dtype, 1234 + rank)
    heat.requires_grad_(True)
    crit = ihpr_b200.JointLocationLoss()

    def step():
        # main/train.py:67-71 (criterion + loss.backward()) as ONE call and ONE launch: loss and d loss / d heat from K5, the gradient
        # handed straight to autograd (heat.grad)
        heat.grad = None
        return crit.forward_backward(heat, gt, vis, hd)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    sampler = ClockSampler(local)           # every rank samples its own GPU: a straggler shows up in sm_mhz_per_rank
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = 0
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        heat.grad = None
        ev[i][0].record()
        crit.forward_backward(heat, gt, vis, hd)
        launches += F.last_launch_count()
        ev[i][1].record()
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t_start.elapsed_time(t_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    launches_per_step = launches / K
    choice = F.last_path_choice() & 3           # 1 = K5 (one launch), 2 = K1 + K2: measured per device on first use (warm-up)

    # ---- the reference's own call sequence through autograd: loss = criterion(...); loss.backward()  (ones-fill + rescale launch extra)
    Ka = min(K, 20)
    for i in range(3 + Ka):
        if i == 3:
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
        heat.grad = None
        crit(heat, gt, vis, hd).backward()
    a1.record()
    torch.cuda.synchronize()
    autograd_ms = a0.elapsed_time(a1) / Ka

    # ---- the two-kernel path (K1 forward, K2 recomputing backward), timed the same way: the standalone rooflines
    crit_u = ihpr_b200.JointLocationLoss(fused_backward=False)
    Ku = min(K, 20)
    evu = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(Ku)]
    for i in range(3 + Ku):
        heat.grad = None
        j = i - 3
        if j >= 0:
            evu[j][0].record()
        l_u = crit_u(heat, gt, vis, hd)
        if j >= 0:
            evu[j][1].record()
        l_u.backward()
        if j >= 0:
            evu[j][2].record()
    torch.cuda.synchronize()
    k1_ms = sum(e[0].elapsed_time(e[1]) for e in evu) / Ku
    k2_ms = sum(e[1].elapsed_time(e[2]) for e in evu) / Ku
    # nvidia-smi samples every 100 ms but the timed region lasts K x 0.3 ms: keep the same step running for another
    # second (untimed) so that the clock / throttle record really is taken under this workload
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.0:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "timed region + 1 s continuation of the same step loop"
    props = torch.cuda.get_device_properties(local)
    per_rank = gather_list({"rank": rank, "launch_ms": fwd_ms, "step_ms": total_ms / K, "k1_ms": k1_ms, "k2_ms": k2_ms, "path_choice": choice,
                            "sms": props.multi_processor_count, "l2_bytes": getattr(props, "L2_cache_size", None),
                            "power_w_max": clocks.get("power_w_max"),
                            "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}, world)
    if world > 1:
        t = torch.tensor([total_ms, fwd_ms, autograd_ms, k1_ms, k2_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, fwd_ms, autograd_ms, k1_ms, k2_ms = t.tolist()
    ms_per_step = total_ms / K
    value = world * R / (ms_per_step * 1e-3)

    # ---- e2e: same step through the host-buffer C-ABI entry point, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h2d_gbs, d2h_gbs, bidir_gbs = measure_pcie(dev)
        hh = torch.empty(heat.shape, dtype=dtype, pin_memory=True)
        hh.copy_(heat.detach())
        out = {"grad": torch.empty(heat.shape, dtype=dtype, pin_memory=True), "loss": torch.empty(1).pin_memory(),
               "coords": torch.empty(B, J, 3).pin_memory()}
        gth, vish, hdh = gt.cpu().pin_memory(), vis.cpu().pin_memory(), hd.cpu().pin_memory()
        ke = args.e2e_steps or min(K, 10)
        for _ in range(2):
            F.integral_l1_fwd_bwd_host(hh, gth, vish, hdh, device=local, slices=args.slices, out=out)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ke):
            F.integral_l1_fwd_bwd_host(hh, gth, vish, hdh, device=local, slices=args.slices, out=out)
        e2e_s = (time.perf_counter() - t0) / ke
        e2e_launch = F.last_launch_count()
        e2e_ranks = gather_list(e2e_s * 1e3, world)
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = t.item()
        small_in = (R * 3 + R + B + 1) * 4
        h2d_b, d2h_b = R * N * es + small_in, R * N * es + R * 3 * 4 + 4
        floor_ms = max(h2d_b / (h2d_gbs * 1e9), d2h_b / (d2h_gbs * 1e9)) * 1e3
        floor_bidir_ms = max(h2d_b, d2h_b) / (bidir_gbs * 1e9) * 1e3
        e2e = {"value": world * R / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_b,
               "d2h_bytes_per_step": d2h_b, "ms_per_step": e2e_s * 1e3, "steps": ke,
               "launches_per_step": e2e_launch, "slices": args.slices,
               "pcie_h2d_GBps": h2d_gbs, "pcie_d2h_GBps": d2h_gbs, "pcie_floor_ms": floor_ms,
               "pcie_bidirectional_GBps_per_direction": bidir_gbs, "pcie_floor_bidirectional_ms": floor_bidir_ms,
               "pcie_floor_note": "bytes / this rank's pinned-copy bandwidth measured alone just before (the two directions overlap: the floor is the "
                                  "slower one); pcie_floor_bidirectional_ms uses the per-direction rate with BOTH directions busy, which is what the "
                                  "pipelined step sees; with N ranks every GPU hangs off the same host memory, so ms_per_step_per_rank rises with N",
               "ms_per_step_per_rank": e2e_ranks,
               "api": "ihpr_integral_l1_fwd_bwd_host (pinned host buffers, synchronous)"}
        del hh, out
        ihpr_b200._lib.lib().ihpr_host_release(local)

    # ---- BASELINE.json configs[1-3]: the training step this path lives in, at the same N (own sub-record)
    train = None
    if not args.no_train and args.dtype == "f32":
        del heat
        torch.cuda.empty_cache()
        try:
            train = train_record(args, dev, world, rank)
        except Exception as e:      # noqa: the headline line must survive a failure of the secondary record
            train = {"error": repr(e)[:300]}

    # ---- the test path (main/test.py:62-65) at test_batch_size and at the training batch, rank 0 only (it does not shard below a batch)
    infer = None
    if not args.no_train and args.dtype == "f32" and rank == 0:
        try:
            infer = infer_record(dev)
        except Exception as e:      # noqa
            infer = {"error": repr(e)[:300]}
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    V = R * N * es
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per launch from the committed ncu capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("fused_ring_kernel_%s_B%d_J%d_D%d" % (args.dtype, B, J, D))
        traffic_src = "constant from the committed ncu --set full capture of this kernel at this shape (%s), not measured in this run" % tj.get("_source", "profiles/")
    if choice == 2:
        # this device measured the two streaming kernels faster than K5: the step is K1 + K2 and the dominant kernel is K2
        roofline = {"bound": "hbm", "kernel": "bwd kernel (K2: re-read h, write dh); the step is K1 + K2 on this device (measured choice)",
                    "achieved": 2 * V / (k2_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                    "algorithmic_bytes_per_launch": 2 * V, "launch_ms": k2_ms, "launch_ms_per_rank": [r["k2_ms"] for r in per_rank]}
        roofline["frac"] = roofline["achieved"] / peak
    roofline_k5 = {"bound": "hbm", "kernel": "fused_ring_kernel (K5: forward + backward in one launch)",
                "achieved": 3 * V / (fwd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 3 * V, "launch_ms": fwd_ms,
                "launch_ms_per_rank": [r["launch_ms"] for r in per_rank],
                "note": "SURVEY 8d counts forward+backward as 3*N*s bytes per joint-volume (read, re-read, write); K5 re-reads from L2, so its "
                        "DRAM traffic is 2*N*s and frac may exceed 1 -- achieved_dram / frac_dram use the 2*N*s it really moves",
                "achieved_dram": 2 * V / (fwd_ms * 1e-3) / 1e9}
    roofline_k5["frac"] = roofline_k5["achieved"] / peak
    roofline_k5["frac_dram"] = roofline_k5["achieved_dram"] / peak
    if choice != 2:
        roofline = roofline_k5
    both = 3 * V / ((k1_ms + k2_ms) * 1e-3) / 1e9
    extra = {"fused_fwd_bwd_ms": fwd_ms, "launches_per_step": launches_per_step,
             "path_choice": {1: "K5 (one launch)", 2: "K1 + K2"}.get(choice, str(choice)),
             "path_choice_note": "variant 0: the first call at this shape on each device timed both forms and kept the faster (include/ihpr_b200.h)",
             "autograd_call_sequence_ms": autograd_ms,
             "autograd_call_sequence_note": "loss = criterion(...); loss.backward() as main/train.py:67-71 writes it: the same K5 launch plus autograd's "
                                            "ones-fill and a rescale launch that exits at once for an upstream gradient of 1",
             "two_kernel_path": {"k1_fwd_ms": k1_ms, "k2_bwd_ms": k2_ms, "k1_fwd_GBps": V / (k1_ms * 1e-3) / 1e9,
                                 "k2_bwd_GBps": 2 * V / (k2_ms * 1e-3) / 1e9, "fwd_bwd_GBps": both,
                                 "fwd_bwd_frac_of_measured": both / peak, "fwd_bwd_frac_of_nominal_8TBps": both / 8000.0,
                                 "volumes_per_s": R / ((k1_ms + k2_ms) * 1e-3)},
             "step_GBps_of_3V_incl_launch_gaps": 3 * V / (ms_per_step * 1e-3) / 1e9,
             "per_rank": per_rank}

    cpu = None
    if not args.no_cpu and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        ref_step, kind, what = load_cpu_reference()
        sB = 8
        v, cores, n, dt = cpu_reference_leg(ref_step, sB, J, D, H, W, min_seconds=10.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "B=%d of the B=%d batch (J=%d, %dx%dx%d fp32), %d fwd+bwd iterations in %.1f s; %s; host has %d logical CPUs"
                         % (sB, B, J, D, H, W, n, dt, what, os.cpu_count())}
        # the reference's own CPU-runnable case (B=1) on ONE thread, reported next to the all-cores figure
        torch.set_num_threads(1)
        v1, _, n1, dt1 = cpu_reference_leg(ref_step, 1, J, D, H, W, min_seconds=4.0, warmup=1)
        torch.set_num_threads(os.cpu_count() or 1)
        cpu["single_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": "B=1 (BASELINE configs[0]), %d iterations in %.1f s" % (n1, dt1)}

    clocks["sm_mhz_per_rank"] = [r["sm_mhz"] for r in per_rank]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": make_config(B, J, D, W, args.dtype),
        "notes": {"residency": "HBM for value / roofline, pinned host memory for e2e", "variant": ihpr_b200.get_variant(),
                  "api": "ihpr_b200.JointLocationLoss().forward_backward(heat, gt, vis, have_depth)  [one launch: K5]"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": extra, "cpu_baseline": cpu,
        "train": train, "infer": infer,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


TRAIN_GFLOP_PER_SAMPLE = {50: 50.55, 152: 108.7}      # conv/deconv MACs x 2, train = 3 x forward, J=18


def EagerJointLocationLoss():
    """What the reference executes on a GPU for its criterion (common/nets/loss.py:13-52): eager ATen ops, autograd
    backward.  Only the `--workload train --torch-loss` comparison arm uses it, to put the e2e gain of the hot path in context."""
    import torch

    class _Eager(torch.nn.Module):
        def forward(self, heat, gt, vis, have_depth):
            J = gt.shape[1]
            B, C, H, W = heat.shape
            D = C // J
            p = torch.softmax(heat.reshape(B, J, D * H * W), 2).reshape(B, J, D, H, W)
            ax, ay, az = p.sum(dim=(2, 3)), p.sum(dim=(2, 4)), p.sum(dim=(3, 4))
            dev = heat.device
            x = (ax * torch.arange(1, W + 1, device=dev, dtype=torch.float32)).sum(2, keepdim=True) - 1
            y = (ay * torch.arange(1, H + 1, device=dev, dtype=torch.float32)).sum(2, keepdim=True) - 1
            z = (az * torch.arange(1, D + 1, device=dev, dtype=torch.float32)).sum(2, keepdim=True) - 1
            loss = torch.abs(torch.cat((x, y, z), 2) - gt) * vis
            return ((loss[:, :, 0] + loss[:, :, 1] + loss[:, :, 2] * have_depth) / 3.).mean()
    return _Eager()


def train_core(args, dev, world, rank, resnet, fused_head, crit_kind, cuda_graph, steps, warmup):
    """BASELINE.json configs[1..3]: ResNet + deconv head + integral loss, synthetic 256x256, B per GPU, one process per GPU, NCCL
    gradient all-reduce (torch DDP).  Backbone / deconvs are stock PyTorch + cuDNN (out of the hot path's scope); final_layer +
    criterion are the sm_100a path.  Returns the record (meaningful on rank 0; timings are the max over ranks)."""
    import types
    import torch
    import torch.distributed as dist
    import ihpr_b200
    from ihpr_b200.model import get_pose_net
    from ihpr_b200.trainer import Trainer, synthetic_batch

    torch.backends.cudnn.benchmark = True                      # main/train.py:34-37
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = args.precision != "fp32"
    B, J = args.batch, args.joints
    cfg = types.SimpleNamespace(resnet_type=resnet, depth_dim=args.depth, input_shape=(4 * args.hw, 4 * args.hw),
                                output_shape=(args.hw, args.hw), lr=1e-3, lr_dec_epoch=[210, 280], lr_dec_factor=0.1, batch_size=B)
    torch.manual_seed(0)
    fused_deconv = bool(fused_head and not getattr(args, "stock_deconv", False))
    net = get_pose_net(cfg, True, J, fused_head=fused_head, fused_deconv=fused_deconv)
    crit = None
    if crit_kind == "torch":
        crit = EagerJointLocationLoss()          # comparison arm only: the reference's eager op sequence on the GPU
    elif crit_kind == "unfused":
        crit = ihpr_b200.JointLocationLoss(fused_backward=False)
    tr = Trainer(net, cfg, criterion=crit, device=dev, autocast_dtype=torch.bfloat16 if args.precision == "bf16" else None,
                 channels_last=True, static_graph=False, graph_capture=cuda_graph)
    host = synthetic_batch(B, J, cfg, None, seed=100 + rank, pin=True)
    devb = [t.to(dev) for t in host]
    if not cuda_graph:          # (with --cuda-graph the warm-up steps run inside capture(), on the stream the graph is captured on)
        for _ in range(max(warmup, 5)):
            tr.train_step(*devb)
    torch.cuda.synchronize()
    K = steps
    step_fn = tr.train_step
    if cuda_graph:
        tr.capture(*devb, warmup=max(warmup, 5))
        step_fn = tr.graphed_step
        for _ in range(3):
            step_fn(*devb)
        torch.cuda.synchronize()

    def timed(from_host, fn=None):
        fn = fn or step_fn
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(K):
            batch = [t.to(dev, non_blocking=True) for t in host] if from_host else devb
            last = fn(*batch)
            if from_host:
                last = last.item()                       # device -> host read of the step's result
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t.tolist()
        return ms / K, wall / K, last

    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_dev, _, _ = timed(False)
    clocks = sampler.stop() if rank == 0 else None
    _, wall_e2e, last = timed(True)
    # what the gradient all-reduce costs on the critical path: the same step with DDP's synchronisation switched off
    exposed = nosync_ms = None
    if world > 1 and not cuda_graph:
        def local_only(*b):
            with tr.model.no_sync():
                return tr.train_step(*b)
        try:
            for _ in range(3):                  # the first steps after switching the synchronisation off are not representative
                local_only(*devb)
            nosync_ms, _, _ = timed(False, local_only)
            sync_again_ms, _, _ = timed(False)  # and the synchronised step once more, right next to it (an eager step is host-paced: drift between
            exposed = min(ms_dev, sync_again_ms) - nosync_ms        # two measurements minutes apart is larger than the quantity)
        except Exception as e:      # noqa
            exposed = "unavailable: %r" % (e,)
    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(peaks_json.get("bf16_tflops_sustained", 1400.0))
    gflop = TRAIN_GFLOP_PER_SAMPLE.get(resnet)
    value = world * B / (ms_dev * 1e-3)
    ach = (gflop * B / (ms_dev * 1e-3) / 1e3) if gflop and J == 18 else None
    h2d = sum(t.numel() * t.element_size() for t in host)
    grad_bytes = sum(p.numel() for p in net.parameters()) * 4
    return {
        "metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": max(warmup, 5),
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "ResNet-%d + deconv head + integral L1 loss training step (Adam), synthetic %dx%d, B=%d per GPU, J=%d, D=%d, "
                               "DDP NCCL all-reduce" % (resnet, 4 * args.hw, 4 * args.hw, B, J, args.depth),
                   "criterion": {"torch": "torch eager (reference ops)", "unfused": "ihpr_b200 K1+K2", "k5": "ihpr_b200 K5 (one launch)",
                                 "head": "ihpr_b200 final_layer + criterion fused on tcgen05 (heat-map never stored)"}[crit_kind],
                   "backbone_head": "stock PyTorch/cuDNN, channels_last, cudnn.benchmark" + (
                       "; deconv blocks 2 / 3 (ConvTranspose2d + BatchNorm + ReLU, forward and backward) on K9 / K10, their weight gradient on cuDNN" if fused_deconv else ""),
                   "launch": "one CUDA graph per step" if cuda_graph else "eager"},
        "clocks": clocks,
        "e2e": {"value": world * B / wall_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": wall_e2e * 1e3,
                "last_loss": last},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": (ach / tf_peak) if ach else None,
                     "note": "whole-step model FLOPs (SURVEY 8d) per GPU / step time vs sustained bf16 GEMM peak; the convolutions are cuDNN's, not this repo's"},
        "allreduce": {"gradient_bytes": grad_bytes, "ms_per_step_without_allreduce": nosync_ms, "exposed_ms": exposed,
                      "note": "exposed = step time with DDP's bucketed NCCL all-reduce (the faster of the timed run and a re-run right after) minus "
                              "the same step under no_sync() (3 warm-up steps first).  A LOWER bound, and it can come out negative on a host-paced eager "
                              "step: under no_sync() autograd ACCUMULATES into DDP's bucket views, one extra add launch per parameter tensor "
                              "(~160 launches, 1-3 ms of host time) that the synchronised step does not have; null at N=1"},
    }


def train_record(args, dev, world, rank):
    """The `train` sub-record of the default bench line: BASELINE.json configs[1] (N=1) / configs[2] (N>1).  The primary numbers are
    the eager step (what main/train.py does: ~1200 launches per step); `cuda_graph` adds the same step replayed as ONE CUDA graph
    (Trainer.capture(): forward, fused-head loss, backward with DDP's bucketed NCCL all-reduce, fused Adam), which is what removes the
    host from the step -- it matters most at N = 8, where eight launch-bound processes share the box's cores."""
    import gc
    import torch
    a = argparse.Namespace(**vars(args))
    a.precision = "bf16"
    rec = train_core(a, dev, world, rank, resnet=50, fused_head=True, crit_kind="head", cuda_graph=False,
                     steps=args.train_steps, warmup=5)
    gc.collect()
    torch.cuda.empty_cache()
    try:
        g = train_core(a, dev, world, rank, resnet=50, fused_head=True, crit_kind="head", cuda_graph=True,
                       steps=args.train_steps, warmup=5)
        rec["cuda_graph"] = {"value": g["value"], "unit": g["unit"], "ms_per_step": g["ms_per_step"],
                             "roofline_frac": g["roofline"]["frac"], "e2e_value": g["e2e"]["value"]}
    except Exception as e:      # noqa: the eager record stands on its own
        rec["cuda_graph"] = {"error": repr(e)[:300]}
    if world == 1:
        # row N1's effect in the driver's own record: the same graphed step with deconv blocks 2 / 3 left to cuDNN + torch BatchNorm
        gc.collect()
        torch.cuda.empty_cache()
        try:
            a2 = argparse.Namespace(**vars(a))
            a2.stock_deconv = True
            s_ = train_core(a2, dev, world, rank, resnet=50, fused_head=True, crit_kind="head", cuda_graph=True,
                            steps=args.train_steps, warmup=5)
            rec["cuda_graph_stock_deconv"] = {"value": s_["value"], "unit": s_["unit"], "ms_per_step": s_["ms_per_step"],
                                              "note": "comparison arm: --stock-deconv (the head's deconv blocks 2 / 3 on cuDNN + nn.BatchNorm2d instead of K9 / K10 / K11)"}
        except Exception as e:      # noqa
            rec["cuda_graph_stock_deconv"] = {"error": repr(e)[:300]}
    g = rec["cuda_graph"]
    rec["best"] = ({"launch": "cuda_graph", "value": g["value"], "ms_per_step": g["ms_per_step"]} if g.get("value", 0) > rec["value"]
                   else {"launch": "eager", "value": rec["value"], "ms_per_step": rec["ms_per_step"]})
    return rec


def infer_record(dev, joint_num=18, depth_dim=64):
    """The `infer` sub-record: ResNet-50 + head -> (B, J, 3) coordinates under eval / no_grad, bf16 autocast for the stock convolutions
    (main/test.py:53-65 with cfg.test_batch_size = 4, main/config.py:44, and at B = 32).  `value` is ResPoseNet(fused_head=True).predict:
    deconv block 3 + BatchNorm + ReLU as K9, final_layer + soft_argmax as K3 (nothing heat-map sized stored); `stock_tail` is the same
    network with the stock module stack producing the heat-map and K1 reading it.  CUDA events around 10 calls each."""
    import types
    import torch
    import ihpr_b200
    from ihpr_b200.model import GraphedPredict, get_pose_net
    cfg = types.SimpleNamespace(resnet_type=50, depth_dim=depth_dim, input_shape=(256, 256), output_shape=(64, 64))
    torch.manual_seed(0)
    net = get_pose_net(cfg, True, joint_num, fused_head=True).to(dev).to(memory_format=torch.channels_last).eval()
    out = {"metric": "inference samples/s", "unit": "samples/s", "data": "synthetic", "dtype": "bf16",
           "config": {"workload": "ResNet-50 + deconv head -> soft-argmax coordinates (eval, no_grad, bf16 autocast), synthetic 256x256, J=%d, D=%d"
                                  % (joint_num, depth_dim)}}

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            fn()
        t1.record()
        torch.cuda.synchronize(dev)
        return t0.elapsed_time(t1) / iters

    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for B in (4, 32):
            img = torch.randn(B, 3, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
            net.fused_head = True
            ms = timed(lambda: net.predict(img))
            c_fused = net.predict(img)
            net.fused_head = False
            ms_stock = timed(lambda: net.predict(img))
            c_stock = net.predict(img)
            out["B%d" % B] = {"value": B / ms * 1e3, "ms_per_batch": ms, "stock_tail": {"value": B / ms_stock * 1e3, "ms_per_batch": ms_stock},
                              "max_coord_diff_vs_stock_tail": float((c_fused - c_stock).abs().max().item())}
            try:        # the same call replayed as ONE CUDA graph (ihpr_b200.model.GraphedPredict): what takes the host out of the test loop
                net.fused_head = True
                gp = GraphedPredict(net, img, autocast_dtype=torch.bfloat16)
                ms_g = timed(lambda: gp(img))
                out["B%d" % B]["cuda_graph"] = {"value": B / ms_g * 1e3, "ms_per_batch": ms_g,
                                                "max_coord_diff_vs_eager": float((gp(img) - c_fused).abs().max().item())}
                del gp
            except Exception as e:      # noqa
                out["B%d" % B]["cuda_graph"] = {"error": repr(e)[:200]}
    del net
    torch.cuda.empty_cache()
    return out


def run_train(args):
    """`--workload train`: the training step as its own JSON line (any ResNet / precision / criterion arm)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload train needs CUDA")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.cuda_graph:
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")        # PyTorch's recipe for capturing DDP's NCCL work
        dist.init_process_group("nccl", device_id=dev)
    kind = "torch" if args.torch_loss else ("unfused" if args.unfused_loss else ("head" if args.fused_head else "k5"))
    rec = train_core(args, dev, world, rank, resnet=args.resnet, fused_head=args.fused_head, crit_kind=kind,
                     cuda_graph=args.cuda_graph, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
