"""bench.py contract checks that need no GPU: the reference arm (CPU oracle on the host cores) prints exactly one JSON line
with the keys the driver reads; under torchrun only rank 0 prints; the product arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
SMALL = ["--workload", "default-640x480-trace", "--steps", "2", "--warmup", "1"]


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", *SMALL], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "default-640x480-trace" and d["config"]["rays_per_frame"] == 548885
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), BENCH, "--impl", "reference", "--gpus", "2", *SMALL]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference" and json.loads(lines[0])["n_gpus"] == 2


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, BENCH, *SMALL], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
