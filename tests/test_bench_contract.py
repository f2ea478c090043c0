"""bench.py contract, the part that runs without a GPU: the reference arm (--impl reference) prints ONE JSON line
with the keys the driver reads, and the b200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    # EVP_B200_LIB points nowhere: the reference arm must not load the product library at all
    env = dict(os.environ, EVP_B200_LIB="/nonexistent/libevp_b200.so")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "square",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "evp_subcycles_per_sec" and d["unit"] == "subcycles/s"
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "square"
    assert set(d["config"]) == {"workload", "cells", "vertices", "active_cells", "active_vertices", "subcycles_per_step",
                                "state", "basis", "l2", "partition", "scaling"}       # the keys of the b200 arm
    assert d["subcycles_per_timed_step"] == 120 and d["extrapolated_from"] is None     # small mesh: whole steps
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "square",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "square"], capture_output=True,
                       text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
