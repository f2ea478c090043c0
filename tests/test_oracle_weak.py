"""CPU known-answer tests of the oracle's weak operators (src/shared/mpas_seaice_velocity_solver_weak.F restated
in oracle/evp_oracle.c): a line integral with edge-midpoint values is exact for linear fields on a planar mesh."""
import numpy as np

import common
import oracle
from mpas_seaice_b200 import variational_init, weakmesh


def _planar():
    mesh, _ = common.mesh_case("hex20")
    return mesh, weakmesh.weak_fields(mesh)


def test_weak_strain_of_a_linear_velocity_field():
    mesh, w = _planar()
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    a, b, c, d = 1.5e-6, -0.7e-6, 0.4e-6, 2.2e-6
    u = np.zeros(nV + 1)
    v = np.zeros(nV + 1)
    u[:nV] = a * mesh.xVertex[:nV] + b * mesh.yVertex[:nV] + 0.3
    v[:nV] = c * mesh.xVertex[:nV] + d * mesh.yVertex[:nV] - 0.1
    e11, e22, e12 = np.zeros(nC + 1), np.zeros(nC + 1), np.zeros(nC + 1)
    solve = np.ones(nC + 1, dtype=np.int32)
    oracle.lib().orc_strain_tensor_weak(
        nC, M, oracle._p(mesh.nEdgesOnCell), oracle._p(mesh.verticesOnCell), oracle._p(mesh.edgesOnCell),
        oracle._p(w["verticesOnEdge"]), oracle._p(mesh.dvEdge), oracle._p(mesh.areaCell), oracle._d(0.0), oracle._p(solve),
        oracle._p(u), oracle._p(v), oracle._p(w["normalVectorPolygon"]), oracle._p(w["latCellRotated"]),
        oracle._p(e11), oracle._p(e22), oracle._p(e12))
    assert np.allclose(e11[:nC], a, rtol=1e-10) and np.allclose(e22[:nC], d, rtol=1e-10)
    assert np.allclose(e12[:nC], 0.5 * (b + c), rtol=1e-10)
    solve[:nC:3] = 0
    oracle.lib().orc_strain_tensor_weak(
        nC, M, oracle._p(mesh.nEdgesOnCell), oracle._p(mesh.verticesOnCell), oracle._p(mesh.edgesOnCell),
        oracle._p(w["verticesOnEdge"]), oracle._p(mesh.dvEdge), oracle._p(mesh.areaCell), oracle._d(0.0), oracle._p(solve),
        oracle._p(u), oracle._p(v), oracle._p(w["normalVectorPolygon"]), oracle._p(w["latCellRotated"]),
        oracle._p(e11), oracle._p(e22), oracle._p(e12))
    assert np.all(e11[:nC:3] == 0.0)


def test_weak_divergence_of_a_linear_stress_field():
    mesh, w = _planar()
    nC, nV, D = mesh.nCells, mesh.nVertices, mesh.vertexDegree
    p, q, r = 2.0e-3, -1.0e-3, 0.5e-3
    s11, s22, s12 = np.zeros(nC + 1), np.zeros(nC + 1), np.zeros(nC + 1)
    s11[:nC] = p * mesh.xCell[:nC] + 7.0
    s22[:nC] = q * mesh.yCell[:nC] - 3.0
    s12[:nC] = r * (mesh.xCell[:nC] + mesh.yCell[:nC])
    interior = variational_init.interior_vertex(mesh)
    sdu, sdv = np.zeros(nV + 1), np.zeros(nV + 1)
    oracle.lib().orc_stress_divergence_weak(
        nV, D, oracle._p(mesh.cellsOnVertex), oracle._p(w["edgesOnVertex"]), oracle._p(mesh.cellsOnEdge),
        oracle._p(mesh.dcEdge), oracle._p(mesh.areaTriangle), oracle._d(0.0), oracle._p(interior),
        oracle._p(s11), oracle._p(s22), oracle._p(s12), oracle._p(w["normalVectorTriangle"]),
        oracle._p(w["latVertexRotated"]), oracle._p(sdu), oracle._p(sdv))
    on = interior[:nV] == 1
    # d(s11)/dx + d(s12)/dy = p + r ;  d(s22)/dy + d(s12)/dx = q + r.  The dual triangle of the generator has
    # areaTriangle = sum of kites, equal to the polygon through the three cell centres on a regular hex mesh.
    assert np.allclose(sdu[:nV][on], p + r, rtol=1e-9) and np.allclose(sdv[:nV][on], q + r, rtol=1e-9)
    assert np.all(sdu[:nV][~on] == 0.0)


def test_weak_and_variational_solutions_agree_qualitatively():
    """The reference's own cross-check of its operators (testing_and_setup/testcases/square/square_quadhex/
    set_difference_fields.py:7-17): same problem, different discretisations, close answers."""
    mesh, var = common.mesh_case("hex20")
    var = dict(var, weak=weakmesh.weak_fields(mesh))
    step, opts = common.step_case(mesh)
    ref = common.run_oracle(mesh, var, step, opts, 120)
    nV = mesh.nVertices
    vm = step["solveVelocity"][:nV] == 1
    scale = np.abs(ref["uVelocity"][:nV][vm]).max()
    for ss, ds in (("weak", "weak"), ("weak", "variational")):
        out = common.run_oracle(mesh, var, step, dict(opts, strain_scheme=ss, stress_divergence_scheme=ds), 120)
        diff = np.abs(out["uVelocity"][:nV][vm] - ref["uVelocity"][:nV][vm])
        assert np.isfinite(out["uVelocity"]).all()
        assert np.median(diff) < 0.05 * scale, (ss, ds, np.median(diff), scale)


def test_final_divergence_shear_weak_uses_the_last_cells_delta():
    """weak.F:729 assigns the whole Delta work array inside the loop; ridgeShear therefore sees the Delta of the
    last owned cell everywhere.  Restated as written."""
    n = 5
    e11 = np.array([1e-7, 2e-7, -3e-7, 0.0, 4e-7])
    e22 = np.array([0.0, -1e-7, 1e-7, 2e-7, -2e-7])
    e12 = np.array([5e-8, 0.0, 0.0, 1e-8, 3e-8])
    div, shear, rc, rs = (np.zeros(n) for _ in range(4))
    oracle.lib().orc_final_divergence_shear_weak(n, oracle._p(e11), oracle._p(e22), oracle._p(e12), oracle._p(div),
                                                 oracle._p(shear), oracle._p(rc), oracle._p(rs))
    assert np.array_equal(div, e11 + e22)
    sd, st, ss = e11[-1] + e22[-1], e11[-1] - e22[-1], 2.0 * e12[-1]
    delta_last = np.sqrt(sd * sd + (st * st + ss * ss) / 4.0)
    assert np.allclose(rs, 0.5 * (delta_last - np.abs(div)), rtol=1e-15)
    assert np.array_equal(rc, -np.minimum(div, 0.0))


def test_final_divergence_shear_weak_against_the_reference_executed_routine():
    """seaice_final_divergence_shear_weak (weak.F:654-751) interpreted from the reference's source
    (tests/golden/options/refexec_weak_post.npz): the oracle reproduces it bit for bit -- including what its
    whole-array assignment `Delta = sqrt(...)` inside the cell loop does to ridgeShear (every cell sees the last owned
    cell's Delta), which the device path restates too (tests/test_gpu_weak.py)."""
    import os
    import oracle
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options", "refexec_weak_post.npz"))
    assert "seaice_final_divergence_shear_weak" in str(z["provenance"])
    nC = int(z["nCells"])
    got = {k: np.zeros(nC + 1) for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear")}
    oracle.lib().orc_final_divergence_shear_weak(nC, oracle._p(z["in_strain11"]), oracle._p(z["in_strain22"]), oracle._p(z["in_strain12"]),
                                                 *[oracle._p(got[k]) for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear")])
    for k in got:
        assert np.array_equal(got[k][:nC], z["out_" + k][:nC]), k
    # the quirk itself: ridgeShear + |divergence| / 2 is the same number in every cell (half the last cell's Delta)
    half_delta = z["out_ridgeShear"][:nC] + 0.5 * np.abs(z["out_divergence"][:nC])
    assert np.ptp(half_delta) <= 1e-21 and half_delta[0] > 0
