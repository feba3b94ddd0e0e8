"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU path timed on the host) prints ONE JSON
line with the keys the driver reads, on the same metric / unit / config.workload naming as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "2", "--joints", "3", "--depth", "8", "--hw", "8")
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"] == "soft-argmax fwd+bwd joint-volumes/s" and d["unit"] == "joint-volumes/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    # the reference itself when baseline/_ref holds its files (build container and GPU box), the oracle port otherwise
    sys.path.insert(0, ROOT)
    from oracle import load_reference
    assert d["cpu_baseline"]["kind"] == ("reference" if load_reference.installed() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0

    # same workload naming as the CUDA arm
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(2, 3, 8, 8, "f32")
    assert d["config"] == bench.make_config(2, 3, 8, 8, "f32")             # the CUDA arm prints exactly this dict too


def test_reference_arm_falls_back_to_the_port_and_says_so(tmp_path, monkeypatch):
    """without baseline/_ref the arm times the oracle port and labels the line accordingly"""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import load_reference
    monkeypatch.setattr(load_reference, "INSTALL_DIR", str(tmp_path / "nothing_here"))
    step, kind, what = bench.load_cpu_reference()
    assert kind == "port" and "port" in what


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--batch", "2", "--joints", "3", "--depth", "8", "--hw", "8"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
