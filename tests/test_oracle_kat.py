"""Pinning the CPU oracle.  The reference stores no numeric outputs and cannot be built here
(SURVEY.md 8c), so the oracle is pinned by the known answers the reference's own test scripts and
code define:

  (1) the analytic operator test of testing_and_setup/testcases/square/operators_strain_stress_divergence
      (fields create_ics.py:12-48, norm strain_stress_divergence_scaling.py:9-27, vertex mask :91-114,
      grid family create_grids.py:181-213): the L2 error must fall between 1st and 2nd order;
  (2) exact reproduction of constant / linear velocity fields by the Wachspress and PWL bases
      (seaice_divergence_stress_test_velocity_set, src/shared/mpas_seaice_testing.F:726-839);
  (3) partition of unity and the 'alternate' denominator identity (variational.F:419-424);
  (4) the closed-form EVP update (constitutive_relation.F:178-248) and the 2x2 solve
      (velocity_solver.F:3172-3203) re-derived independently in numpy for single points.
"""
import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import meshgen, synthetic


def _slot_mask(mesh):
    return np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:mesh.nCells, None]


def _operator_setup(mesh, u, v, cr="linear"):
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    interior = synthetic.interior_vertex(mesh)
    z = lambda: np.zeros(nV + 1)
    zc = lambda: np.zeros((nC + 1, M))
    ss = np.ones(nC + 1, dtype=np.int32)
    ss[nC] = 0
    step = dict(solveStress=ss, solveVelocity=interior.copy(), icePressure=np.zeros(nC + 1),
                uVelocity=u.copy(), vVelocity=v.copy(), stress11=zc(), stress22=zc(), stress12=zc(),
                strain11=zc(), strain22=zc(), strain12=zc(), replacementPressure=zc(),
                totalMassVertex=z(), totalMassVertexfVertex=z(), iceAreaVertex=z(), airStressVertexU=z(),
                airStressVertexV=z(), surfaceTiltForceU=z(), surfaceTiltForceV=z(), oceanStressU=z(),
                oceanStressV=z(), uOceanVelocityVertex=z(), vOceanVelocityVertex=z(), stressDivergenceU=z(),
                stressDivergenceV=z(), oceanStressCoeff=z(), uVelocityInitial=z(), vVelocityInitial=z())
    opts = dict(constitutive_relation_type=cr, ocean_stress_type="quadratic", use_ocean_stress=True,
                elasticTimeStep=30.0, dynamicsTimeStep=3600.0, dampingTimescale=1296.0)
    return step, opts


def _use_vertex(mesh):
    """get_use_vertex (strain_stress_divergence_scaling.py:91-114): drop every vertex of every cell that
    is a neighbour of a non-interior cell."""
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    coc = mesh.cellsOnCell[:nC]
    n = mesh.nEdgesOnCell[:nC]
    slot = np.arange(M)[None, :] < n[:, None]
    interior_cell = np.all(~slot | (coc <= nC), axis=1)
    use = np.ones(nV, dtype=bool)
    for c in np.nonzero(~interior_cell)[0]:
        for k in range(n[c]):
            c2 = coc[c, k] - 1
            if c2 < nC:
                use[mesh.verticesOnCell[c2, :n[c2]] - 1] = False
    return use


def _l2(num, ana, area, use):
    return np.sqrt(np.sum(area[use] * (num[use] - ana[use]) ** 2) / np.sum(area[use] * ana[use] ** 2))


@pytest.mark.parametrize("basis", ["wachspress", "pwl"])
def test_operator_convergence_hex(basis):
    """BASELINE configs[0]: strain_stress_divergence operator test, planar hex 82x94 family, variational
    scheme, linear constitutive relation, one subcycle, single CPU rank."""
    errs, res = [], []
    for nx, ny in ((42, 48), (82, 94), (162, 186)):
        dc = 1.0 / (nx - 2)
        mesh = meshgen.planar_hex(nx, ny, dc)
        var = oracle.init_variational(mesh, basis=basis, metric=False)
        ana = synthetic.operator_test_fields(mesh)
        step, opts = _operator_setup(mesh, ana["u"], ana["v"])
        oracle.subcycle_velocity_solver(mesh, var, step, opts, 1)
        nV = mesh.nVertices
        use = _use_vertex(mesh) & (step["solveVelocity"][:nV] == 1)
        assert use.sum() > 0.7 * nV
        area = mesh.areaTriangle[:nV]
        errs.append((_l2(step["stressDivergenceU"][:nV], ana["divu"][:nV], area, use),
                     _l2(step["stressDivergenceV"][:nV], ana["divv"][:nV], area, use)))
        res.append(dc)
        # linear relation: sigma == epsilon (constitutive_relation.F:366-371); no velocity update (:2529-2541)
        sm = _slot_mask(mesh)
        assert np.array_equal(step["stress11"][:mesh.nCells][sm], step["strain11"][:mesh.nCells][sm])
        assert np.array_equal(step["uVelocity"], ana["u"])
        # strain at the cell-vertex stress points against the analytic strain at that vertex
        voc = mesh.verticesOnCell[:mesh.nCells] - 1
        e11 = step["strain11"][:mesh.nCells]
        rel = np.abs(e11[sm] - ana["e11"][voc[sm]]).max() / np.abs(ana["e11"]).max()
        assert rel < 0.3 * (dc / 0.025)      # first order at the one-sided stress points
    errs = np.array(errs)
    order = np.log2(errs[:-1] / errs[1:])
    assert np.all(errs[-1] < 2e-2), errs
    assert np.all(order > 0.9) and np.all(order < 2.6), (errs, order)


@pytest.mark.parametrize("kind,basis", [("hex20", "wachspress"), ("hex20", "pwl"), ("quad40", "wachspress"),
                                        ("quad40", "pwl"), ("ico3", "wachspress"), ("ico3", "pwl")])
def test_basis_reproduces_constant_and_linear_fields(kind, basis):
    mesh, _ = common.mesh_case(kind)
    var = oracle.init_variational(mesh, basis=basis, metric=False)
    nC, M = mesh.nCells, mesh.maxEdges
    GU, GV = var["basisGradientU"][:nC], var["basisGradientV"][:nC]     # [c, j(gradient vertex), i(basis)]
    xl, yl = var["xLocal"][:nC], var["yLocal"][:nC]
    sm = _slot_mask(mesh)
    scale = np.abs(GU).max()
    # sum_i grad phi_i = 0 at every stress point; sum_i x_i dphi_i/dx = 1, sum_i x_i dphi_i/dy = 0 ...
    assert np.abs(GU.sum(axis=2)[sm]).max() < 1e-11 * scale
    assert np.abs(GV.sum(axis=2)[sm]).max() < 1e-11 * scale
    assert np.abs(np.einsum("cji,ci->cj", GU, xl)[sm] - 1.0).max() < 1e-10
    assert np.abs(np.einsum("cji,ci->cj", GV, yl)[sm] - 1.0).max() < 1e-10
    assert np.abs(np.einsum("cji,ci->cj", GU, yl)[sm]).max() < 1e-10
    assert np.abs(np.einsum("cji,ci->cj", GV, xl)[sm]).max() < 1e-10
    if not mesh.on_a_sphere:
        # u = x, v = -y (testing.F:726-839 style linear fields): e11 = 1, e22 = -1, e12 = 0 to round-off
        step, opts = _operator_setup(mesh, mesh.xVertex.copy(), -mesh.yVertex.copy())
        oracle.subcycle_velocity_solver(mesh, var, step, opts, 1)
        assert np.abs(step["strain11"][:nC][sm] - 1.0).max() < 1e-9
        assert np.abs(step["strain22"][:nC][sm] + 1.0).max() < 1e-9
        assert np.abs(step["strain12"][:nC][sm]).max() < 1e-9
        # constant stress => zero divergence at vertices whose cells are all interior to the support
        nV = mesh.nVertices
        use = _use_vertex(mesh) & (step["solveVelocity"][:nV] == 1)
        scale_div = 1.0 / mesh.dc
        assert np.abs(step["stressDivergenceU"][:nV][use]).max() < 1e-8 * scale_div
        assert np.abs(step["stressDivergenceV"][:nV][use]).max() < 1e-8 * scale_div
    # integrals: sum_i of int(phi_i dphi_j) = int(dphi_j) and sum_j of that = 0 (partition of unity);
    # sum_ij int(phi_i phi_j) = cell area in the local tangent plane
    SU, SV, SM = var["basisIntegralsU"][:nC], var["basisIntegralsV"][:nC], var["basisIntegralsMetric"][:nC]
    s_scale = np.abs(SU).max()
    assert np.abs(SU.sum(axis=(1, 2))).max() < 1e-10 * s_scale * M
    assert np.abs(SV.sum(axis=(1, 2))).max() < 1e-10 * s_scale * M
    n = mesh.nEdgesOnCell[:nC]
    poly = np.zeros(nC)
    for s in range(M):
        ok = n > s
        nxt = np.where(s + 1 < n, s + 1, 0)
        poly[ok] += 0.5 * (xl[ok, s] * yl[np.nonzero(ok)[0], nxt[ok]] - xl[np.nonzero(ok)[0], nxt[ok]] * yl[ok, s])
    # PWL rescales its sub-triangles to the SPHERICAL areaCell (pwl.F:159-183): differs from the tangent-plane
    # polygon by O((dc/R)^2) on the coarse test sphere
    tol = 1e-2 if (mesh.on_a_sphere and basis == "pwl") else 1e-9
    assert np.abs(SM.sum(axis=(1, 2)) / poly - 1.0).max() < tol


def test_wachspress_gradient_sparsity_and_quadrature():
    """Gradients are stored only for iGradientVertex in {i-1, i, i+1} (wachspress.F:1178-1191); Dunavant
    order 8 has 16 points whose weights sum to 1 and norm 2 (wachspress.F:1573-1597, :1426)."""
    mesh, var = common.mesh_case("ico3")
    nC, M = mesh.nCells, mesh.maxEdges
    GU = var["basisGradientU"][:nC]
    n = mesh.nEdgesOnCell[:nC]
    for c in (0, 5, nC - 1, int(np.nonzero(n == 5)[0][0])):
        for i in range(n[c]):
            for j in range(n[c]):
                near = (j - i) % n[c] in (0, 1, n[c] - 1)
                if not near:
                    assert GU[c, j, i] == 0.0
    u, v, w, norm = oracle.integration_factors("dunavant", 8)
    assert len(w) == 16 and norm == 2.0
    assert abs(w.sum() - 1.0) < 1e-12
    assert np.all(u >= 0) and np.all(v >= 0) and np.all(u + v <= 1.0 + 1e-14)
    # exactness on the unit triangle: int x^a y^b = a! b! / (a+b+2)!
    from math import factorial
    for a, b in ((1, 0), (2, 1), (3, 3), (8, 0), (4, 4)):
        exact = factorial(a) * factorial(b) / factorial(a + b + 2)
        assert abs(np.sum(w * u ** a * v ** b) / norm - exact) < 2e-13, (a, b)


def test_alternate_denominator_identity():
    """'alternate' = sum of basisIntegralsMetric over the stress points of the cells at the vertex
    (variational.F:403-443); on a regular planar hex mesh it equals areaTriangle to round-off."""
    mesh, _ = common.mesh_case("hex20")
    var = oracle.init_variational(mesh, denominator="alternate", metric=False)
    interior = synthetic.interior_vertex(mesh)[:mesh.nVertices] == 1
    den = var["variationalDenominator"][:mesh.nVertices]
    assert np.abs(den[interior] / mesh.areaTriangle[:mesh.nVertices][interior] - 1.0).max() < 1e-10


def test_evp_point_update_against_closed_form():
    """One stress point, one subcycle: constitutive_relation.F:178-248 re-derived in numpy."""
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh)
    rng = np.random.default_rng(7)
    nV = mesh.nVertices
    step["uVelocity"][:nV] = 0.1 * rng.standard_normal(nV)
    step["vVelocity"][:nV] = 0.1 * rng.standard_normal(nV)
    s0 = {k: 100.0 * rng.standard_normal(step[k].shape) for k in ("stress11", "stress22", "stress12")}
    for k in s0:
        step[k][:] = s0[k]
        step[k][step["solveStress"] != 1] = 0.0
        s0[k] = step[k].copy()
    out = common.run_oracle(mesh, var, step, opts, 1)
    cm, vm = common.masks_for(mesh, step)
    e11, e22, e12 = out["strain11"], out["strain22"], out["strain12"]
    # strain itself from the gradients (variational.F:633-668), planar => no metric
    voc = mesh.verticesOnCell - 1
    u = np.append(step["uVelocity"], 0.0)[np.minimum(voc, nV)]
    v = np.append(step["vVelocity"], 0.0)[np.minimum(voc, nV)]
    GU, GV = var["basisGradientU"], var["basisGradientV"]
    assert np.allclose(np.einsum("cji,ci->cj", GU, u)[cm], e11[cm], rtol=1e-12, atol=1e-18)
    assert np.allclose(np.einsum("cji,ci->cj", 0.5 * GV, u)[cm] + np.einsum("cji,ci->cj", 0.5 * GU, v)[cm], e12[cm],
                       rtol=1e-11, atol=1e-18)
    dte, T = opts["elasticTimeStep"], opts["dampingTimescale"]
    P = step["icePressure"][:, None]
    sd, st, ss = e11 + e22, e11 - e22, 2.0 * e12
    Delta = np.sqrt(sd ** 2 + (st ** 2 + ss ** 2) / 4.0)
    pc = P / np.maximum(Delta, 1e-11)
    rep = pc * Delta
    pc = pc * dte / (2.0 * T)
    den = 1.0 + 0.5 * dte / T
    s1 = (s0["stress11"] + s0["stress22"] + pc * (sd - Delta)) / den
    s2 = (s0["stress11"] - s0["stress22"] + pc / 4.0 * st) / den
    s12 = (s0["stress12"] + pc / 4.0 * ss * 0.5) / den
    assert np.allclose(out["replacementPressure"][cm], rep[cm], rtol=1e-13)
    assert np.allclose(out["stress11"][cm], (0.5 * (s1 + s2))[cm], rtol=1e-12, atol=1e-9)
    assert np.allclose(out["stress22"][cm], (0.5 * (s1 - s2))[cm], rtol=1e-12, atol=1e-9)
    assert np.allclose(out["stress12"][cm], s12[cm], rtol=1e-12, atol=1e-9)
    # 2x2 solve (velocity_solver.F:3172-3203): residual of the linear system it claims to solve
    C = out["oceanStressCoeff"]
    m, mf = step["totalMassVertex"], step["totalMassVertexfVertex"]
    un, vn = out["uVelocity"], out["vVelocity"]
    r1 = (m / dte + C) * un - mf * vn - (out["stressDivergenceU"] + step["airStressVertexU"] + step["surfaceTiltForceU"]
                                        + C * step["oceanStressU"] + m * step["uVelocity"] / dte)
    r2 = (m / dte + C) * vn + mf * un - (out["stressDivergenceV"] + step["airStressVertexV"] + step["surfaceTiltForceV"]
                                        + C * step["oceanStressV"] + m * step["vVelocity"] / dte)
    scale = np.abs(m[vm] / dte * un[vm]).max()
    assert np.abs(r1[vm]).max() < 1e-11 * scale and np.abs(r2[vm]).max() < 1e-11 * scale
    dragio, rhow = 0.00536, 1026.0
    Cref = dragio * rhow * step["iceAreaVertex"] * np.sqrt((step["uOceanVelocityVertex"] - step["uVelocity"]) ** 2 +
                                                          (step["vOceanVelocityVertex"] - step["vVelocity"]) ** 2)
    assert np.allclose(C[vm], Cref[vm], rtol=1e-13)


def test_integer_maps_against_python_loops():
    """cellVerticesAtVertex / interiorVertex (mesh.F:632-685, 423-488): bit-exact integer work, checked
    against a literal pure-Python restatement on small meshes."""
    for kind in ("hex20", "quad40", "ico3"):
        mesh, var = common.mesh_case(kind)
        nC, nV, D = mesh.nCells, mesh.nVertices, mesh.vertexDegree
        cv = np.zeros((nV + 1, D), dtype=np.int32)
        it = np.zeros(nV + 1, dtype=np.int32)
        for iv in range(nV):
            ok = 0
            for k in range(D):
                c = mesh.cellsOnVertex[iv, k]
                if 1 <= c <= nC:
                    ok += 1
                for j in range(mesh.nEdgesOnCell[c - 1]):
                    if mesh.verticesOnCell[c - 1, j] == iv + 1:
                        cv[iv, k] = j + 1
            it[iv] = int(ok == D)
        assert np.array_equal(cv, var["cellVerticesAtVertex"])
        assert np.array_equal(it, var["interiorVertex"])
