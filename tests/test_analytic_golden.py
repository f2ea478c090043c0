"""Known answers computed by the REFERENCE'S OWN test-case scripts (tests/golden/analytic_*.npz, generated in the
build container by tests/golden/make_analytic_golden.py importing /root/reference/testing_and_setup/testcases/...):
the analytic velocity, strain and stress-divergence fields of its operator tests
  - planar:  square/operators_strain_stress_divergence/create_ics.py:12-48   (BASELINE configs[0], hex 82 x 94)
  - sphere:  spherical_operators/strain_stress_divergence/create_ic.py:574-603 (10 242 cells, rotated unit sphere)
with the error norm of strain_stress_divergence_scaling.py:9-27 and its near-boundary vertex mask (:91-114).
The reference encodes no pass/fail threshold (it plots error against resolution between first- and second-order
guide lines); the thresholds below are the measured errors of the restated operators plus 25 % head-room, i.e.
they pin today's behaviour against the reference's analytic truth.
CPU: the oracle.  GPU: the device gives the oracle's bits, hence the same norms."""
import os

import numpy as np
import pytest

import oracle
from mpas_seaice_b200 import meshgen, synthetic
from test_oracle_kat import _operator_setup, _slot_mask, _use_vertex

HERE = os.path.dirname(os.path.abspath(__file__))


def l2_norm(numerical, analytical, area, use):
    """L2_norm of strain_stress_divergence_scaling.py:9-27 (area-weighted, relative)."""
    return float(np.sqrt(np.sum(area[use] * (numerical[use] - analytical[use]) ** 2) / np.sum(area[use] * analytical[use] ** 2)))


def _planar_case():
    g = np.load(os.path.join(HERE, "golden", "analytic_planar_hex82.npz"))
    mesh = meshgen.planar_hex(int(g["nx"]), int(g["ny"]), float(g["dc"]))
    nV = mesh.nVertices
    assert np.array_equal(mesh.xVertex[:nV], g["x"]) and np.array_equal(mesh.yVertex[:nV], g["y"])
    var = oracle.init_variational(mesh, metric=False)
    u, v = np.zeros(nV + 1), np.zeros(nV + 1)
    u[:nV], v[:nV] = g["u"], g["v"]
    step, opts = _operator_setup(mesh, u, v)
    use = _use_vertex(mesh) & (step["solveVelocity"][:nV] == 1)
    return g, mesh, var, step, opts, use


def _sphere_case():
    g = np.load(os.path.join(HERE, "golden", "analytic_sphere_ico5.npz"))
    mesh = meshgen.icosphere(int(g["level"]), radius=1.0)
    nV = mesh.nVertices
    assert np.array_equal(mesh.xVertex[:nV], g["x"]) and np.array_equal(mesh.zVertex[:nV], g["z"])
    var = oracle.init_variational(mesh)          # rotated grid + metric terms, Registry defaults
    u, v = np.zeros(nV + 1), np.zeros(nV + 1)
    u[:nV], v[:nV] = g["u"], g["v"]
    step, opts = _operator_setup(mesh, u, v)
    return g, mesh, var, step, opts, np.ones(nV, dtype=bool)


# measured (oracle, this repo): planar divu 1.38e-2, divv 1.24e-2, strains 8.3e-2 .. 1.0e-1 (first order at the
# one-sided stress points); sphere divu 4.2e-2, divv 3.7e-2, strains 5.7e-2 .. 7.2e-2
PLANAR_LIMITS = dict(divu=0.0175, divv=0.0155, e11=0.126, e22=0.104, e12=0.116)
SPHERE_LIMITS = dict(divu=0.053, divv=0.047, e11=0.080, e22=0.072, e12=0.090)


def _norms(g, mesh, step, use):
    nC, nV = mesh.nCells, mesh.nVertices
    area = mesh.areaTriangle[:nV]
    out = dict(divu=l2_norm(step["stressDivergenceU"][:nV], g["divu"], area, use),
               divv=l2_norm(step["stressDivergenceV"][:nV], g["divv"], area, use))
    sm = _slot_mask(mesh)
    voc = mesh.verticesOnCell[:nC] - 1
    cell_use = use[np.where(sm, voc, 0)] & sm                 # stress points whose vertex is used
    w = np.broadcast_to(mesh.areaCell[:nC, None], sm.shape)
    for k, name in (("e11", "strain11"), ("e22", "strain22"), ("e12", "strain12")):
        num = step[name][:nC]
        ana = g[k][np.where(sm, voc, 0)]
        out[k] = float(np.sqrt(np.sum(w[cell_use] * (num[cell_use] - ana[cell_use]) ** 2) / np.sum(w[cell_use] * ana[cell_use] ** 2)))
    return out


def test_reference_planar_fields_equal_our_restatement():
    """synthetic.operator_test_fields restates create_ics.py:12-48; the reference's own output agrees to round-off."""
    g, mesh, *_ = _planar_case()
    ana = synthetic.operator_test_fields(mesh)
    nV = mesh.nVertices
    for k in ("u", "v", "e11", "e22", "e12", "divu", "divv"):
        assert np.allclose(ana[k][:nV], g[k], rtol=1e-12, atol=1e-12 * np.abs(g[k]).max()), k


@pytest.mark.parametrize("case,limits", [("planar", PLANAR_LIMITS), ("sphere", SPHERE_LIMITS)])
def test_oracle_operators_against_reference_analytic_fields(case, limits):
    g, mesh, var, step, opts, use = _planar_case() if case == "planar" else _sphere_case()
    oracle.subcycle_velocity_solver(mesh, var, step, opts, 1)
    norms = _norms(g, mesh, step, use)
    for k, lim in limits.items():
        assert norms[k] < lim, (k, norms)
    assert use.sum() > 0.7 * mesh.nVertices


@pytest.mark.gpu
@pytest.mark.parametrize("case,limits", [("planar", PLANAR_LIMITS), ("sphere", SPHERE_LIMITS)])
def test_device_operators_against_reference_analytic_fields(evp_lib, case, limits):
    import common
    g, mesh, var, step, opts, use = _planar_case() if case == "planar" else _sphere_case()
    out = common.run_device(mesh, var, step, opts, 1)
    norms = _norms(g, mesh, out, use)
    for k, lim in limits.items():
        assert norms[k] < lim, (k, norms)
    ref = common.run_oracle(mesh, var, step, opts, 1)
    for k in ("strain11", "strain22", "strain12", "stressDivergenceU", "stressDivergenceV"):
        assert np.array_equal(out[k], ref[k]), k


# measured (oracle, this repo): planar divu 3.7e-2, divv 2.6e-2, strains 3.4e-3 (second order at the cell centre);
# sphere divu 3.1e-2, divv 3.0e-2, strains 3.1e-3 .. 3.4e-3
WEAK_LIMITS = {"planar": dict(divu=0.046, divv=0.033, strain=0.0043), "sphere": dict(divu=0.039, divv=0.038, strain=0.0042)}


@pytest.mark.parametrize("case", ["planar", "sphere"])
def test_weak_operators_against_reference_analytic_fields(case):
    """The 'weak' variant of the reference's operator tests (run_model.py:16 lists wachspress, pwl, weak, ...): weak
    strain at the cell centres and weak stress divergence at the vertices against the same analytic fields.  On the
    sphere this also validates mpas_seaice_b200.weakmesh's restatement of the reference's normal vectors
    (mesh.F:1038-1606 with removeMetricTerms = .true.): with per-edge frames instead, strain22 is off by v tan(lat)/R."""
    from mpas_seaice_b200 import weakmesh
    g, mesh, var, step, opts, use = _planar_case() if case == "planar" else _sphere_case()
    wk = weakmesh.weak_fields(mesh)
    for k in ("normalVectorPolygon", "normalVectorTriangle", "latCellRotated", "latVertexRotated"):
        assert np.isfinite(wk[k]).all(), k
    o = dict(opts, strain_scheme="weak", stress_divergence_scheme="weak")
    oracle.subcycle_velocity_solver(mesh, dict(var, weak=wk), step, o, 1)
    nC, nV = mesh.nCells, mesh.nVertices
    lim = WEAK_LIMITS[case]
    area = mesh.areaTriangle[:nV]
    assert l2_norm(step["stressDivergenceU"][:nV], g["divu"], area, use) < lim["divu"]
    assert l2_norm(step["stressDivergenceV"][:nV], g["divv"], area, use) < lim["divv"]
    sm = _slot_mask(mesh)
    voc = np.where(sm, mesh.verticesOnCell[:nC] - 1, 0)
    n = mesh.nEdgesOnCell[:nC]
    cell_use = np.all(use[voc] | ~sm, axis=1)
    w = mesh.areaCell[:nC]
    for k, name in (("e11", "strain11Weak"), ("e22", "strain22Weak"), ("e12", "strain12Weak")):
        ana = np.where(sm, g[k][voc], 0.0).sum(axis=1) / n          # analytic strain averaged over the cell's vertices
        num = step[name][:nC]
        err = float(np.sqrt(np.sum(w[cell_use] * (num[cell_use] - ana[cell_use]) ** 2) / np.sum(w[cell_use] * ana[cell_use] ** 2)))
        assert err < lim["strain"], (k, err)


# the seven operator variants of the reference's own test scripts (square/operators_strain_stress_divergence/
# run_model.py:16, spherical_operators/strain_stress_divergence/run_model.py:17):
#   name: (basis, strain scheme, stress divergence scheme, average variational strain)
VARIANTS = {"wachspress": ("wachspress", "variational", "variational", False),
            "pwl": ("pwl", "variational", "variational", False),
            "weak": ("wachspress", "weak", "weak", False),
            "wachsavg": ("wachspress", "variational", "variational", True),
            "pwlavg": ("pwl", "variational", "variational", True),
            "weakwachs": ("wachspress", "weak", "variational", False),
            "weakpwl": ("pwl", "weak", "variational", False)}
# measured relative L2 errors of (stressDivergenceU, stressDivergenceV), oracle, this repo; the test allows +25 %
MEASURED = {"planar": {"wachspress": (0.0138, 0.0124), "pwl": (0.0227, 0.0137), "weak": (0.0368, 0.0265),
                       "wachsavg": (0.0125, 0.0114), "pwlavg": (0.0145, 0.0145), "weakwachs": (0.0180, 0.0179),
                       "weakpwl": (0.0178, 0.0178)},
            "sphere": {"wachspress": (0.0420, 0.0372), "pwl": (0.0414, 0.0366), "weak": (0.0307, 0.0299),
                       "wachsavg": (0.0399, 0.0326), "pwlavg": (0.0398, 0.0329), "weakwachs": (0.0423, 0.0334),
                       "weakpwl": (0.0421, 0.0335)}}


@pytest.mark.parametrize("case", ["planar", "sphere"])
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_all_seven_operator_variants_against_reference_analytic_fields(case, variant):
    from mpas_seaice_b200 import weakmesh
    basis, ss, ds, avg = VARIANTS[variant]
    g, mesh, var, step, opts, use = _planar_case() if case == "planar" else _sphere_case()
    if basis == "pwl":
        var = oracle.init_variational(mesh, basis="pwl", metric=(False if case == "planar" else None))
    var = dict(var, weak=weakmesh.weak_fields(mesh))
    o = dict(opts, strain_scheme=ss, stress_divergence_scheme=ds, average_variational_strain=avg)
    oracle.subcycle_velocity_solver(mesh, var, step, o, 1)
    nV = mesh.nVertices
    area = mesh.areaTriangle[:nV]
    eu = l2_norm(step["stressDivergenceU"][:nV], g["divu"], area, use)
    ev = l2_norm(step["stressDivergenceV"][:nV], g["divv"], area, use)
    mu, mv = MEASURED[case][variant]
    assert eu < 1.25 * mu and ev < 1.25 * mv, (eu, ev)
    assert eu > 0.5 * mu and ev > 0.5 * mv, (eu, ev)          # a sudden improvement is a change of behaviour too
