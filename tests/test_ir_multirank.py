"""N > 1 for the transport: one process per rank (gloo), each with its block (two halo layers), geometry and
transport through the C ABI, tracer halo update over the process group after every step.  Owned cells must be
bit-identical to the single-block oracle run.  CPU: the kernels under host emulation, world sizes 2 and 3.
GPU: libir_b200.so, one device per rank (needs >= 2 devices; not yet run on a device)."""
import os
import socket
import subprocess
import sys
import time

import numpy as np
import pytest

from oracle import ir
from mpas_seaice_b200 import ir_host
from test_oracle_ir import case, smooth_divergent_velocity, _random_state
from test_ir_parity import _emulation_library, clone

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(lib_path, kind, n_steps, world, out_path, timeout=200):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs, logs = [], []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2", IR_RANK_TRACE_AFTER=str(timeout - 20))
        log = open(out_path + f".rank{r}.log", "wb")
        logs.append(log)
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_ir_rank_worker.py"), lib_path, kind, str(n_steps),
                                       out_path], env=env, stdout=log, stderr=subprocess.STDOUT))
    t_end = time.time() + timeout
    try:
        while time.time() < t_end and any(p.poll() is None for p in procs):
            if any(p.poll() not in (None, 0) for p in procs):
                break
            time.sleep(0.2)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
                p.wait()
        for log in logs:
            log.close()
    outs = [open(out_path + f".rank{r}.log", errors="replace").read() for r in range(world)]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} rc={p.returncode}:\n" + "\n".join(
            f"--- rank {q} ---\n{o[-3000:]}" for q, o in enumerate(outs))
    return dict(np.load(out_path))


def _single_block(kind, n_steps):
    mesh, irf, geom = case(kind)
    rng = np.random.default_rng(17)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    ref = clone(tracers)
    for _ in range(n_steps):
        ir.run(mesh, irf, geom, ref, u, v, 3600.0)
    return mesh, ref


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_reproduce_the_single_block_run_under_emulation(world, tmp_path):
    mesh, ref = _single_block("ico4", 3)
    out = _launch(_emulation_library(), "ico4", 3, world, str(tmp_path / "ir.npz"))
    for t in ref:
        assert np.array_equal(out[t.name][:mesh.nCells], t.array[:mesh.nCells]), t.name


@pytest.mark.gpu
def test_ranks_reproduce_the_single_block_run_on_devices(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    mesh, ref = _single_block("ico4", 3)
    out = _launch(ir_host.LIB_PATH, "ico4", 3, 2, str(tmp_path / "ir.npz"))
    for t in ref:
        assert np.array_equal(out[t.name][:mesh.nCells], t.array[:mesh.nCells]), t.name
