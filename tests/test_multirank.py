"""N > 1: the decomposed path end to end, one process per rank.

CPU (gloo, world_size 2 and 3): host logic -- per-rank block construction, request exchange through the
host communicator, the evp_set_halo lists driving real point-to-point messages -- with the oracle as
block solver; owned results must be BIT-identical to the single-rank oracle run.
GPU (nccl, needs >= 2 devices): the same through libevp_b200.so and its in-graph NCCL exchange."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import common
from mpas_seaice_b200 import workloads

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode, name, nsub, world, out_path, method="auto", timeout=240):
    """One process per rank; each rank's output goes to a file so that a hang can be diagnosed (the
    workers dump their Python stacks shortly before the timeout)."""
    port = _free_port()
    procs, logs = [], []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2", EVP_RANK_TRACE_AFTER=str(max(10, timeout - 20)))
        log = open(out_path + f".rank{r}.log", "wb")
        logs.append(log)
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_rank_worker.py"), mode, name, str(nsub),
                                       out_path, method], env=env, stdout=log, stderr=subprocess.STDOUT))
    import time
    t_end = time.time() + timeout
    try:
        while time.time() < t_end and any(p.poll() is None for p in procs):
            if any(p.poll() not in (None, 0) for p in procs):
                break                       # one rank failed: do not wait for the others to time out
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


def _reference(name, nsub):
    import oracle
    w = workloads.build(name, with_static=False)
    mesh, step, opts = w["mesh"], w["step"], w["opts"]
    var = oracle.init_variational(mesh)
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    return mesh, step, ref


def _diff_report(k, a, b):
    bad = np.nonzero(a != b)[0]
    rel = np.abs(a[bad] - b[bad]) / np.maximum(np.abs(b[bad]), 1e-300)
    return (f"{k}: {bad.size} of {a.size} entries differ, max rel diff {rel.max():.3e}, first flat indices "
            f"{bad[:8].tolist()}, got {a[bad[:3]].tolist()} expected {b[bad[:3]].tolist()}")


def _assert_equal(mesh, step, ref, out):
    cm, vm = common.masks_for(mesh, step)
    problems = []
    for k in common.COMPARE_CELL:
        if not np.array_equal(out[k][cm], ref[k][cm]):
            problems.append(_diff_report(k, out[k][cm], ref[k][cm]))
    for k in common.COMPARE_VERTEX:
        if not np.array_equal(out[k][vm], ref[k][vm]):
            problems.append(_diff_report(k, out[k][vm], ref[k][vm]))
    assert not problems, "\n".join(problems)
    assert np.abs(ref["uVelocity"]).max() > 0


@pytest.mark.parametrize("world,method", [(2, "auto"), (3, "rcb")])
def test_gloo_ranks_match_single_rank(tmp_path, world, method):
    name, nsub = "ico3", 6
    out = _launch("oracle", name, nsub, world, str(tmp_path / "out.npz"), method)
    mesh, step, ref = _reference(name, nsub)
    _assert_equal(mesh, step, ref, out)


@pytest.mark.gpu
def test_nccl_full_dynamics_step_matches_single_rank(evp_lib, tmp_path):
    """pre-subcycle + 120 subcycles + post-subcycle on the device, decomposed (2 halo layers, ice caps):
    the u,v exchange that closes new_ice_velocities and the oceanStress exchange of ocean_stress_final
    (velocity_solver.F:1281-1320, 3735-3760) ride on the same NCCL lists."""
    import torch
    import oracle
    from mpas_seaice_b200 import synthetic, variational_init
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    name, nsub = "ico4", 120
    out = _launch("gpu-full", name, nsub, world, str(tmp_path / "out.npz"), timeout=150)
    w = workloads.build(name, state="B", with_static=False)
    mesh, opts = w["mesh"], w["opts"]
    var = oracle.init_variational(mesh)
    state = synthetic.sphere_state(mesh, "B")
    ref = oracle.pre_subcycle(mesh, state, w["config_dt"])
    oracle.subcycle_velocity_solver(mesh, var, ref, opts, nsub)
    interior = variational_init.interior_vertex(mesh)
    ref.update(oracle.final_divergence_shear(mesh, ref))
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, ref, opts, interior)
    post_ref = dict(ref, oceanStressU=osu, oceanStressV=osv, oceanStressCellU=ocu, oceanStressCellV=ocv)
    nC, nV = mesh.nCells, mesh.nVertices
    cm, vm = common.masks_for(mesh, ref)
    problems = []
    for k in ("stress11", "stress22", "stress12"):
        if not np.array_equal(out[k][cm], ref[k][cm]):
            problems.append(_diff_report(k, out[k][cm], ref[k][cm]))
    for k in ("uVelocity", "vVelocity") + common.POST_VERTEX:
        if not np.array_equal(out[k][vm], post_ref[k][vm]):
            problems.append(_diff_report(k, out[k][vm], post_ref[k][vm]))
    for k in common.POST_CELL:
        if not np.array_equal(out[k][:nC], post_ref[k][:nC]):
            problems.append(_diff_report(k, out[k][:nC], post_ref[k][:nC]))
    assert not problems, "\n".join(problems)
    assert np.abs(post_ref["oceanStressCellU"]).max() > 0 and 0 < cm.sum() < cm.size


@pytest.mark.gpu
@pytest.mark.parametrize("name,nsub,mode", [("ico4", 120, "gpu-p2p"), ("square", 120, "gpu-p2p"), ("ico4", 120, "gpu-nccl"),
                                            ("ico4", 7, "gpu-p2p"), ("ico3", 120, "gpu")])
def test_nccl_ranks_match_single_rank(evp_lib, tmp_path, name, nsub, mode):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    out = _launch(mode, name, nsub, world, str(tmp_path / "out.npz"), timeout=150)
    mesh, step, ref = _reference(name, nsub)
    _assert_equal(mesh, step, ref, out)
