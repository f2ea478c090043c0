"""examples/minimal_host.c: the C ABI used from plain C.

CPU: the example compiles against include/evp_b200.h as C99 with -Wall -Wextra, links to the shipped library,
fails loudly without a device (no CPU fallback), and the mesh/basis arrays it builds are the same the oracle's
Wachspress precompute gives for that mesh -- so the example's arrays follow the MPAS layout the ABI documents.
GPU: the example runs and prints the expected strain.
"""
import ctypes as C
import os
import subprocess
import types

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "minimal_host.c")
LIBDIR = os.path.join(ROOT, "mpas-seaice_b200", "csrc")


def _build(tmp_path, shared=False):
    out = str(tmp_path / ("minimal_host.so" if shared else "minimal_host"))
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
           "-L" + LIBDIR, "-levp_b200", "-Wl,-rpath," + LIBDIR, "-o", out]
    if shared:
        cmd[1:1] = ["-shared", "-fPIC"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def _fill(tmp_path):
    L = C.CDLL(_build(tmp_path, shared=True))
    nC, nV, M, D = 9, 16, 4, 4
    a = dict(nEdgesOnCell=np.zeros(nC + 1, np.int32), verticesOnCell=np.zeros((nC + 1, M), np.int32),
             cellsOnVertex=np.zeros((nV + 1, D), np.int32), cellVerticesAtVertex=np.zeros((nV + 1, D), np.int32),
             basisGradientU=np.zeros((nC + 1, M, M)), basisGradientV=np.zeros((nC + 1, M, M)),
             basisIntegralsU=np.zeros((nC + 1, M, M)), basisIntegralsV=np.zeros((nC + 1, M, M)),
             basisIntegralsMetric=np.zeros((nC + 1, M, M)), tanLatVertexRotatedOverRadius=np.zeros(nV + 1),
             variationalDenominator=np.zeros(nV + 1))
    L.example_fill.restype = None
    L.example_fill(*[C.c_void_p(v.ctypes.data) for v in a.values()])
    return a, (nC, nV, M, D)


def test_example_fails_loudly_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3
    assert "evp_create failed" in r.stderr


def test_example_arrays_give_the_expected_strain_in_the_oracle(tmp_path):
    """One linear-rheology subcycle of the oracle on the example's arrays: u = 1e-6 x -> strain11 = 1e-6."""
    a, (nC, nV, M, D) = _fill(tmp_path)
    assert (a["nEdgesOnCell"][:nC] == 4).all()
    # every interior vertex has four cells, and cellVerticesAtVertex inverts verticesOnCell
    for v in range(nV):
        for s in range(D):
            c = a["cellsOnVertex"][v, s]
            if c <= nC:
                assert a["verticesOnCell"][c - 1, a["cellVerticesAtVertex"][v, s] - 1] == v + 1
    # a gradient basis sums to zero over the basis functions, the mass integrals sum to the cell area
    assert np.abs(a["basisGradientU"][:nC].sum(axis=2)).max() < 1e-18      # numpy (cell, grad vertex, basis vertex)
    assert np.allclose(a["basisIntegralsMetric"][:nC].sum(axis=(1, 2)), 1000.0 ** 2)
    mesh = types.SimpleNamespace(nCells=nC, nVertices=nV, maxEdges=M, vertexDegree=D, on_a_sphere=False)
    var = dict(a, areaCell=np.full(nC + 1, 1000.0 ** 2))
    ix, iy = np.arange(nV) % 4, np.arange(nV) // 4
    zc, zv = np.zeros(nC + 1), np.zeros(nV + 1)
    step = dict(solveStress=np.r_[np.ones(nC, np.int32), np.int32(0)],
                solveVelocity=np.r_[((ix > 0) & (ix < 3) & (iy > 0) & (iy < 3)).astype(np.int32), np.int32(0)],
                icePressure=zc.copy(), uVelocity=np.r_[1.0e-6 * 1000.0 * ix, 0.0], vVelocity=zv.copy())
    for k in ("totalMassVertex", "totalMassVertexfVertex", "iceAreaVertex", "airStressVertexU", "airStressVertexV",
              "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "oceanStressV", "uOceanVelocityVertex",
              "vOceanVelocityVertex", "stressDivergenceU", "stressDivergenceV", "oceanStressCoeff"):
        step[k] = zv.copy()
    step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
    for k in ("stress11", "stress22", "stress12", "strain11", "strain22", "strain12"):
        step[k] = np.zeros((nC + 1, M))
    step["replacementPressure"] = np.zeros((nC + 1, M))
    opts = dict(constitutive_relation_type="linear", elasticTimeStep=30.0, dynamicsTimeStep=3600.0,
                dampingTimescale=1296.0)

    class _M(dict):
        __getattr__ = dict.__getitem__
    m = _M(vars(mesh))
    oracle.subcycle_velocity_solver(m, var, step, opts, 1)
    assert np.allclose(step["strain11"][:nC], 1.0e-6, rtol=1e-12)
    assert np.abs(step["strain22"][:nC]).max() < 1e-18 and np.abs(step["strain12"][:nC]).max() < 1e-18


def test_example_basis_equals_the_wachspress_precompute(tmp_path):
    """On a square the Wachspress functions are the bilinear ones the example writes down in closed form."""
    from mpas_seaice_b200 import meshgen
    a, (nC, nV, M, D) = _fill(tmp_path)
    mesh = meshgen.planar_quad(5, 5, 1000.0)
    var = oracle.init_variational(mesh)
    # an interior cell of the reference mesh, its vertices reordered to the example's (-,-) (+,-) (+,+) (-,+)
    c = int(np.argmin(np.hypot(mesh.xCell[:mesh.nCells] - mesh.xCell[:mesh.nCells].mean(),
                               mesh.yCell[:mesh.nCells] - mesh.yCell[:mesh.nCells].mean())))
    vs = mesh.verticesOnCell[c, :4] - 1
    dx, dy = mesh.xVertex[vs] - mesh.xCell[c], mesh.yVertex[vs] - mesh.yCell[c]
    order = [int(np.nonzero((np.sign(dx) == sx) & (np.sign(dy) == sy))[0][0])
             for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1))]
    for name in ("basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric"):
        ref = var[name][c][np.ix_(order, order)]
        got = a[name][4]
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 2e-12 * scale, name


@pytest.mark.gpu
def test_example_runs_on_the_device(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "strain11 at cell 5, vertex 1: 1.000000e-06" in r.stdout
