"""The init-time precompute against golden vectors produced by EXECUTING THE REFERENCE'S FORTRAN SOURCE
(tests/golden/init/refexec_init_*.npz, made by tests/golden/make_reference_executed_golden.py with the interpreter
tests/golden/fortran_subset.py): init_velocity_solver_variational_primary_mesh (variational.F:108-344) with the Wachspress
(wachspress.F:46-1287) or piecewise-linear (pwl.F:44-373, numerics.F LU) basis, the local coordinates, the metric terms,
cellVerticesAtVertex and the variational denominator.

CPU: oracle/evp_precompute_oracle.c reproduces all eight arrays bit for bit.  GPU: so does evp_precompute_wachspress /
evp_precompute_pwl through the C ABI.  Tolerance: none (the reference's statements, one IEEE operation each, libm for the
trigonometry of the sphere)."""
import ast
import glob
import os

import numpy as np
import pytest

import oracle
from mpas_seaice_b200 import meshgen

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "init", "refexec_init_*.npz")))
# generated after this round's GPU time was spent (the 'fekete' rules, dunavant order 12): oracle replay only
CPU_FILES = sorted(glob.glob(os.path.join(HERE, "golden", "cpu", "refexec_init_*.npz")))
OUT = ("cellVerticesAtVertex", "tanLatVertexRotatedOverRadius", "basisGradientU", "basisGradientV", "basisIntegralsU",
       "basisIntegralsV", "basisIntegralsMetric", "variationalDenominator")


def _load(path):
    z = np.load(path)
    spec = ast.literal_eval(str(z["spec"]))
    mesh = getattr(meshgen, spec[0])(*spec[1:])
    for k in z.files:
        if k.startswith("mesh_"):           # the generated mesh is the one the reference's statements saw
            assert np.array_equal(mesh[k[5:]], z[k]), "meshgen no longer produces the fixture's mesh: " + k
    want = {k: z["out_" + k] for k in OUT}
    kw = dict(basis=str(z["basis"]), denominator=str(z["denominator"]), integration_type=str(z["integration_type"]),
              integration_order=int(z["integration_order"]))
    return mesh, want, kw, str(z["provenance"])


def test_fixtures_exist_and_cover_both_bases():
    assert len(FILES) >= 5
    seen = set()
    for f in FILES:
        prov = _load(f)[3]
        assert "interpreting the reference's Fortran source" in prov
        seen |= {w.strip() for w in prov.split(":", 1)[1].split(",")}
    for name in ("init_velocity_solver_variational_primary_mesh", "seaice_init_velocity_solver_wachspress",
                 "integrate_wachspress_polygon", "wachspress_basis_derivative", "get_integration_factors_dunavant",
                 "get_integration_factors_trapezoidal", "seaice_init_velocity_solver_pwl", "lu_decomposition",
                 "calc_local_coords_spherical", "calc_local_coords_planar", "seaice_calc_variational_metric_terms",
                 "seaice_cell_vertices_at_vertex", "variational_denominator"):
        assert name in seen, name


@pytest.mark.parametrize("path", FILES + CPU_FILES, ids=[os.path.basename(f)[13:-4] for f in FILES + CPU_FILES])
def test_oracle_precompute_reproduces_the_reference_executed_arrays(path):
    mesh, want, kw, _ = _load(path)
    var = oracle.init_variational(mesh, **kw)
    for k in OUT:
        assert np.array_equal(var[k], want[k]), k
    assert np.abs(want["basisIntegralsU"]).max() > 0 and np.abs(want["basisGradientU"]).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[13:-4] for f in FILES])
def test_device_precompute_reproduces_the_reference_executed_arrays(evp_lib, path):
    import common
    from mpas_seaice_b200 import host
    mesh, want, kw, _ = _load(path)
    var = oracle.init_variational(mesh, **kw)            # xLocal / yLocal and the host-side maps the create call takes
    step, opts = common.step_case(mesh)
    extra = dict(integration=(kw["integration_type"], kw["integration_order"])) if kw["basis"] == "wachspress" else dict(basis="pwl")
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]), **extra)
    try:
        got = solver.fetch_basis()
    finally:
        solver.destroy()
    nC = mesh.nCells
    for k, a in got.items():
        assert np.array_equal(a[:nC], want[k][:nC]), k
