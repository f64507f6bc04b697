"""bench.py contract on a CPU-only box: the reference arm prints ONE JSON line with the required keys, and the
default arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-rays", "30000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["steps"] == 2 and j["value"] > 1e5
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "workload" in j["config"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]


def test_reference_arm_under_torchrun_env_uses_all_cores_and_stays_bounded():
    """torch.distributed.run exports OMP_NUM_THREADS=1: rank 0 must still use every core it may run on, and the K steps must
    fit the time budget whatever K is (round 1: 20 steps x 118 s on one thread -> the driver's 870 s limit, no N>1 ratio)."""
    import time
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="8", LOCAL_RANK="0")
    t0 = time.time()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--steps", "20",
                          "--warmup", "3", "--ref-budget", "8"], capture_output=True, text=True, timeout=200, env=env)
    wall = time.time() - t0
    assert out.returncode == 0, out.stderr[-400:]
    j = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count()
    assert j["cpu_baseline"]["cores"] == avail and j["n_gpus"] == 8 and j["steps"] == 20
    assert j["scaling"] == "strong" and j["config"]["rays_per_step"] == 1_000_000_000
    assert j["ms_per_step"] * 20 < 3 * 8e3 and wall < 90, (j["ms_per_step"], wall)
    if avail > 1:      # more than one thread really ran: well above the ~1e7 bounces/s of one core
        assert j["value"] > 1.3e7, j["value"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--ref-rays", "1000"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--rays", "1000"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
