"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
stored inputs -> stored outputs.  CPU: the oracle still reproduces them bit for bit (pins the checker
against drift between rounds).  GPU: libevp_b200.so reproduces them bit for bit through the C-ABI."""
import glob
import os

import numpy as np
import pytest

import common
from mpas_seaice_b200 import meshgen

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz"))
               if not os.path.basename(f).startswith("analytic_"))      # those belong to test_analytic_golden.py


def _load(path):
    z = np.load(path)
    mesh = meshgen.Mesh()
    var, step, opts, out = {}, {}, {}, {}
    for k in z.files:
        a = z[k]
        a = a.item() if a.ndim == 0 else a
        if k.startswith("mesh_"):
            mesh[k[5:]] = a
        elif k.startswith("var_"):
            var[k[4:]] = a
        elif k.startswith("in_"):
            step[k[3:]] = a
        elif k.startswith("opt_"):
            opts[k[4:]] = a
        elif k.startswith("out_"):
            out[k[4:]] = a
    return mesh, var, step, opts, out, int(z["nsub"])


def _check(mesh, step, want, got):
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_CELL:
        assert np.array_equal(got[k][cm], want[k][cm]), k
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(got[k][vm], want[k][vm]), k


def test_golden_files_exist():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    mesh, var, step, opts, want, nsub = _load(path)
    got = common.run_oracle(mesh, var, step, opts, nsub)
    _check(mesh, step, want, got)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_device_reproduces_golden(evp_lib, path):
    from mpas_seaice_b200 import host
    mesh, var, step, opts, want, nsub = _load(path)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        if opts.get("average_variational_strain"):
            interior = np.zeros(mesh.nVertices + 1, dtype=np.int32)          # not read by the averaging
            ext = meshgen.Mesh(mesh)
            ext["cellsOnCell"] = np.zeros((mesh.nCells + 1, mesh.maxEdges), dtype=np.int32)
            ext["areaTriangle"] = np.ones(mesh.nVertices + 1)
            ext["fVertex"] = np.zeros(mesh.nVertices + 1)
            solver.set_mesh_ext(ext, interior)
        solver.update_step(step)
        solver.run_subcycles(nsub)
        got = solver.fetch()
    finally:
        solver.destroy()
    _check(mesh, step, want, got)
