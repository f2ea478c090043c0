"""Committed golden vectors (tests/golden/*.npz): stored inputs -> stored outputs, replayed bit for bit through the
oracle (CPU) and through libevp_b200.so and the C ABI (GPU).

* refexec_*.npz -- outputs computed by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCE (subcycle_velocity_solver and
  everything below it) with the interpreter tests/golden/fortran_subset.py; made by
  tests/golden/make_reference_executed_golden.py.  These pin the oracle -- and the device -- to the reference itself.
* the others -- made by tests/golden/make_golden.py from the oracle: longer runs on larger meshes, a guard against drift
  between rounds."""
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
    var, step, opts, out, weak = {}, {}, {}, {}, {}
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
        elif k.startswith("weak_"):
            weak[k[5:]] = a
    if weak:                                   # the velocity_weak pool's static arrays (weak operators)
        var["weak"] = weak
        nC, nV = int(mesh["nCells"]), int(mesh["nVertices"])
        for k in WEAK_STATE:
            step[k] = np.zeros((nV if k.endswith("Vertex") else nC) + 1)
    return mesh, var, step, opts, out, int(z["nsub"])


WEAK_STATE = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak",
              "replacementPressureWeak", "strain11Vertex", "strain22Vertex", "strain12Vertex")


def _check(mesh, step, want, got, opts=None):
    cm, vm = common.masks_for(mesh, step)
    opts = opts or {}
    weak_strain = opts.get("strain_scheme", "variational") == "weak"
    weak_div = opts.get("stress_divergence_scheme", "variational") == "weak"
    if weak_strain:
        nC = mesh.nCells
        for k in (WEAK_STATE[:7] if weak_div else WEAK_STATE[3:6]):
            assert np.array_equal(got[k][:nC], want[k][:nC]), k
        assert np.abs(want["strain11Weak"]).max() > 0
    for k in (() if weak_div else common.COMPARE_CELL):
        assert np.array_equal(got[k][cm], want[k][cm]), k
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(got[k][vm], want[k][vm]), k
    if "vertexBoundaryType" in step:          # the special-boundary vertices themselves (not solved: outside vm)
        b = step["vertexBoundaryType"] != 0
        assert b.any() and np.array_equal(got["uVelocity"][b], want["uVelocity"][b]) and np.array_equal(got["vVelocity"][b], want["vVelocity"][b])


def test_golden_files_exist():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    mesh, var, step, opts, want, nsub = _load(path)
    got = common.run_oracle(mesh, var, step, opts, nsub)
    _check(mesh, step, want, got, opts)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_device_reproduces_golden(evp_lib, path):
    from mpas_seaice_b200 import host
    mesh, var, step, opts, want, nsub = _load(path)
    sb = (step["vertexBoundaryType"], step["vertexBoundarySourceLocal"]) if opts.get("use_special_boundaries_velocity") else None
    solver = host.EvpSolver(mesh, var, opts, special_boundaries=sb)
    try:
        if opts.get("average_variational_strain"):
            interior = np.zeros(mesh.nVertices + 1, dtype=np.int32)          # not read by the averaging
            ext = meshgen.Mesh(mesh)
            ext["cellsOnCell"] = np.zeros((mesh.nCells + 1, mesh.maxEdges), dtype=np.int32)
            ext["areaTriangle"] = np.ones(mesh.nVertices + 1)
            ext["fVertex"] = np.zeros(mesh.nVertices + 1)
            solver.set_mesh_ext(ext, interior)
        solver.update_step(step)
        if "weak" in var:
            solver.set_weak_mesh(mesh, var["weak"])
            solver.update_weak_state({k: step[k] for k in WEAK_STATE[:3]})
        solver.run_subcycles(nsub)
        got = solver.fetch()
        if "weak" in var:
            got.update(solver.fetch_weak())
    finally:
        solver.destroy()
    _check(mesh, step, want, got, opts)


CPU_FILES = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "cpu", "refexec_*.npz"))
                   if not any(w in f for w in ("_step_", "_init_", "locked_cells", "special_boundaries_init", "quadrature_rules")))   # (those: test_refexec_step.py, test_host_numpy.py)


@pytest.mark.parametrize("path", CPU_FILES, ids=[os.path.basename(f)[:-4] for f in CPU_FILES])
def test_oracle_reproduces_reference_executed_vectors_added_late(path):
    """Reference-executed vectors generated after the round's GPU time was spent (piecewise-linear basis with the
    'alternate' denominator, no ocean stress, constitutive relation 'none'): replayed through the oracle here; the device
    runs the same configurations against the oracle in tests/test_gpu_parity.py (test_device_pwl_precompute_bit_exact,
    test_no_ocean_stress, test_namelist_options)."""
    assert len(CPU_FILES) == 4
    mesh, var, step, opts, want, nsub = _load(path)
    got = common.run_oracle(mesh, var, step, opts, nsub)
    if opts.get("use_special_boundaries_velocity_masks"):      # the masks the subcycle ran with are the special-boundary ones
        step = dict(step, solveStress=step["solveStressSpecialBoundaries"], solveVelocity=step["solveVelocitySpecialBoundaries"])
    _check(mesh, step, want, got, opts)


def test_oracle_made_vectors_equal_their_reference_executed_twins():
    """hex20_evp_120 / ico3_revised_40 / ico3_evp_120 exist twice: made by the oracle (make_golden.py) and made by
    interpreting the reference's source from the same inputs at the same length (a whole 120-subcycle dynamics step of the
    square case and of the sphere).  Same inputs, same outputs on every solved cell and vertex, bit for bit."""
    pairs = [(f, os.path.join(HERE, "golden", "refexec_" + os.path.basename(f))) for f in FILES
             if not os.path.basename(f).startswith("refexec_")]
    pairs = [(a, b) for a, b in pairs if os.path.exists(b)]
    assert len(pairs) >= 2
    for a, b in pairs:
        mesh, var, step, opts, want, nsub = _load(a)
        mesh2, var2, step2, opts2, want2, nsub2 = _load(b)
        assert nsub == nsub2 and "interpreting the reference's Fortran source" in str(np.load(b)["provenance"])
        for k in ("uVelocity", "icePressure", "totalMassVertex", "solveStress", "solveVelocity"):
            assert np.array_equal(step[k], step2[k]), k
        for k in ("basisGradientU", "basisIntegralsMetric"):
            assert np.array_equal(var[k], var2[k]), k
        _check(mesh, step, want, want2, opts)
