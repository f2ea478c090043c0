"""Known answers that do NOT come from the oracle's own C: closed-form solutions of the EVP stress update and of
the vertex momentum solve iterated over a whole dynamics step (120 subcycles), evaluated here in numpy / Python
complex arithmetic from the formulas of the reference, and held against BOTH the oracle (CPU) and the device path
(GPU).  With the reference's analytic operator fields (test_oracle_kat.py, test_analytic_golden.py) this is as far
as the oracle can be pinned without a Fortran compiler: the strain / divergence operators by the reference's own
analytic test, the EVP branch and the 2x2 solve by their exact recurrences.

(1) Uniform strain rate, velocities held fixed (no vertex is solved).  A linear velocity field is reproduced exactly
    by the Wachspress basis, so every stress point sees the same (e11, e22, e12) and the EVP update
    (src/shared/mpas_seaice_velocity_solver_constitutive_relation.F:225-245)
        s1 <- (s1 + c (eD - Delta)) / d,  s2 <- (s2 + c eT / e^2) / d,  s12 <- (s12 + c eS / (2 e^2)) / d,
        c = (P / max(Delta, puny)) dte / (2 T),  d = 1 + dte / (2 T)
    is a geometric sequence with ratio 1/d towards the viscous-plastic fixed point
        s1* = P (eD - Delta) / Delta,  s2* = P eT / (e^2 Delta),  s12* = P eS / (2 e^2 Delta).
(2) No internal stress (P = 0), uniform forcing: the vertex solve (velocity_solver.F:3172-3203) with linear ocean
    drag is, in w = u + i v,
        w <- (tau + C w_o + (m / dte) w) / (m / dte + C + i m f),   C = dragio rhow a,
    a geometric sequence with complex ratio q = (m/dte) / (m/dte + C + i m f) towards (tau + C w_o) / (C + i m f).
"""
import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import synthetic

N_SUB = 120
DTE, T_DAMP = 30.0, 1296.0
ECC2 = 4.0            # eccentricity squared, constitutive_relation.F:41-43
PUNY = 1.0e-11
DRAGIO, RHOW = 0.00536, 1026.0      # ice_constants_colpkg.F90


def _blank_step(mesh, u, v):
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    z = lambda: np.zeros(nV + 1)
    zc = lambda: np.zeros((nC + 1, M))
    ss = np.ones(nC + 1, dtype=np.int32)
    ss[nC] = 0
    return dict(solveStress=ss, solveVelocity=np.zeros(nV + 1, dtype=np.int32), icePressure=np.zeros(nC + 1),
                uVelocity=u.copy(), vVelocity=v.copy(), stress11=zc(), stress22=zc(), stress12=zc(),
                strain11=zc(), strain22=zc(), strain12=zc(), replacementPressure=zc(),
                totalMassVertex=z(), totalMassVertexfVertex=z(), iceAreaVertex=z(), airStressVertexU=z(),
                airStressVertexV=z(), surfaceTiltForceU=z(), surfaceTiltForceV=z(), oceanStressU=z(),
                oceanStressV=z(), uOceanVelocityVertex=z(), vOceanVelocityVertex=z(), stressDivergenceU=z(),
                stressDivergenceV=z(), oceanStressCoeff=z(), uVelocityInitial=z(), vVelocityInitial=z())


def _opts(**kw):
    o = dict(constitutive_relation_type="evp", ocean_stress_type="quadratic", use_ocean_stress=True,
             elasticTimeStep=DTE, dynamicsTimeStep=3600.0, dampingTimescale=T_DAMP)
    o.update(kw)
    return o


# ------------------------------------------------------------------------------------------------------------------
# (1) EVP relaxation towards the viscous-plastic state under a uniform strain rate
# ------------------------------------------------------------------------------------------------------------------
STRAINS = [(1.0e-6, -0.4e-6, 0.3e-6, 0.5e-6),      # a, b, c, d of u = a x + b y, v = c x + d y: shear + convergence
           (-2.0e-7, 0.0, 0.0, -1.0e-7),            # pure convergence
           (0.0, 3.0e-7, -3.0e-7, 0.0)]             # rigid rotation: zero strain, Delta = 0 -> the puny branch


def _evp_case(kind, coeffs):
    mesh, var = common.mesh_case(kind, metric=False)
    a, b, c, d = coeffs
    x, y = mesh.xVertex, mesh.yVertex
    step = _blank_step(mesh, a * x + b * y, c * x + d * y)
    nC = mesh.nCells
    P = 2.75e4 * 2.0 * np.exp(-20.0 * (1.0 - 0.9))
    step["icePressure"][:nC] = P
    s0 = (1.3e3, -0.7e3, 0.4e3)                      # start away from zero so that the decay term is seen too
    step["stress11"][:nC], step["stress22"][:nC], step["stress12"][:nC] = s0
    return mesh, var, step, P, s0


def _evp_closed_form(coeffs, P, s0, n):
    a, b, c, d = coeffs
    e11, e22, e12 = a, d, 0.5 * (b + c)
    eD, eT, eS = e11 + e22, e11 - e22, 2.0 * e12
    Delta = np.sqrt(eD * eD + (eT * eT + eS * eS) / ECC2)
    pOverDelta = P / max(Delta, PUNY)
    fixed1 = pOverDelta * (eD - Delta)
    fixed2 = pOverDelta * eT / ECC2
    fixed12 = pOverDelta * eS * 0.5 / ECC2
    r = (1.0 / (1.0 + 0.5 * DTE / T_DAMP)) ** n
    s1 = fixed1 + (s0[0] + s0[1] - fixed1) * r
    s2 = fixed2 + (s0[0] - s0[1] - fixed2) * r
    s12 = fixed12 + (s0[2] - fixed12) * r
    return 0.5 * (s1 + s2), 0.5 * (s1 - s2), s12, pOverDelta * Delta


def _check_evp(mesh, out, coeffs, P, s0):
    nC = mesh.nCells
    sm = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
    x11, x22, x12, rep = _evp_closed_form(coeffs, P, s0, N_SUB)
    scale = max(abs(x11), abs(x22), abs(x12))
    # the strain of a linear field is exact up to round-off of the basis gradients (1e-10 relative, test_oracle_kat.py),
    # which the Delta-normalised update passes on to the stresses
    for name, want in (("stress11", x11), ("stress22", x22), ("stress12", x12)):
        got = out[name][:nC][sm]
        assert np.abs(got - want).max() <= 2e-8 * scale, (name, float(np.abs(got - want).max() / scale))
    # Delta = 0 analytically (rigid rotation): the computed Delta is round-off, which P / puny magnifies by 1e11
    assert np.abs(out["replacementPressure"][:nC][sm] - rep).max() <= (1e-9 * rep if rep > 0.0 else 1e-8 * P)
    # velocities were not to be touched
    assert np.array_equal(out["uVelocity"][:mesh.nVertices], (coeffs[0] * mesh.xVertex + coeffs[1] * mesh.yVertex)[:mesh.nVertices])


@pytest.mark.parametrize("kind", ["hex20", "quad40"])
@pytest.mark.parametrize("coeffs", STRAINS)
def test_evp_relaxation_oracle(kind, coeffs):
    mesh, var, step, P, s0 = _evp_case(kind, coeffs)
    out = common.run_oracle(mesh, var, step, _opts(), N_SUB)
    _check_evp(mesh, out, coeffs, P, s0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex20", "quad40"])
@pytest.mark.parametrize("coeffs", STRAINS)
def test_evp_relaxation_device(evp_lib, kind, coeffs):
    mesh, var, step, P, s0 = _evp_case(kind, coeffs)
    out = common.run_device(mesh, var, step, _opts(), N_SUB)
    _check_evp(mesh, out, coeffs, P, s0)


def test_evp_fixed_point_is_the_viscous_plastic_state():
    """After many damping time scales the EVP stress IS the VP stress of Hibler's elliptical yield curve: on the
    ellipse ((s1 + P)/P)^2 + e^2 (s2^2 + 4 s12^2)/P^2 = 1 for Delta > puny (a property the closed form cannot hide)."""
    coeffs = STRAINS[0]
    mesh, var, step, P, s0 = _evp_case("hex20", coeffs)
    out = common.run_oracle(mesh, var, step, _opts(), 4000)          # (1/d)^4000 ~ 1e-20
    nC = mesh.nCells
    sm = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
    s1 = (out["stress11"] + out["stress22"])[:nC][sm]
    s2 = (out["stress11"] - out["stress22"])[:nC][sm]
    s12 = out["stress12"][:nC][sm]
    # s1 = P (eD - Delta)/Delta  =>  (s1 + P)/P = eD/Delta; s2 = P eT/(e^2 Delta); 2 s12 = P eS/(e^2 Delta)
    ellipse = ((s1 + P) / P) ** 2 + ECC2 * (s2 ** 2 + 4.0 * s12 ** 2) / P ** 2
    assert np.abs(ellipse - 1.0).max() < 1e-7


# ------------------------------------------------------------------------------------------------------------------
# (2) vertex momentum solve: wind against Coriolis and linear ocean drag, no internal stress
# ------------------------------------------------------------------------------------------------------------------
def _momentum_case(kind, use_ocean):
    mesh, var = common.mesh_case(kind, metric=False)
    nV = mesh.nVertices
    w0 = 0.03 - 0.02j
    step = _blank_step(mesh, np.full(nV + 1, w0.real), np.full(nV + 1, w0.imag))
    step["solveVelocity"] = synthetic.interior_vertex(mesh).astype(np.int32)
    m, f, area = 917.0 * 1.9, 1.46e-4, 0.93
    tau, wo = 0.11 + 0.04j, 0.08 - 0.05j
    step["totalMassVertex"][:] = m
    step["totalMassVertexfVertex"][:] = m * f
    step["iceAreaVertex"][:] = area
    step["airStressVertexU"][:] = tau.real
    step["airStressVertexV"][:] = tau.imag
    step["uOceanVelocityVertex"][:] = wo.real
    step["vOceanVelocityVertex"][:] = wo.imag
    step["oceanStressU"][:] = wo.real        # ocean_stress with turning angle 0 (velocity_solver.F:1846-1878)
    step["oceanStressV"][:] = wo.imag
    C = DRAGIO * RHOW * area if use_ocean else 0.0
    q = (m / DTE) / (m / DTE + C + 1j * m * f)
    fixed = (tau + C * wo) / (C + 1j * m * f)
    want = fixed + (w0 - fixed) * q ** N_SUB
    return mesh, var, step, want, w0


def _check_momentum(mesh, step, out, want, w0):
    nV = mesh.nVertices
    solved = step["solveVelocity"][:nV] == 1
    assert solved.sum() > 0.5 * nV
    got = out["uVelocity"][:nV] + 1j * out["vVelocity"][:nV]
    assert np.abs(got[solved] - want).max() <= 1e-12 * abs(want)
    assert np.all(got[~solved] == w0)                       # boundary vertices are not solved
    assert abs(want - w0) > 0.1 * abs(w0)                   # and the test did move the ice


@pytest.mark.parametrize("kind", ["hex20", "quad40"])
@pytest.mark.parametrize("use_ocean", [True, False])
def test_momentum_solve_oracle(kind, use_ocean):
    mesh, var, step, want, w0 = _momentum_case(kind, use_ocean)
    out = common.run_oracle(mesh, var, step, _opts(ocean_stress_type="linear", use_ocean_stress=use_ocean), N_SUB)
    _check_momentum(mesh, step, out, want, w0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex20", "quad40"])
@pytest.mark.parametrize("use_ocean", [True, False])
def test_momentum_solve_device(evp_lib, kind, use_ocean):
    mesh, var, step, want, w0 = _momentum_case(kind, use_ocean)
    out = common.run_device(mesh, var, step, _opts(ocean_stress_type="linear", use_ocean_stress=use_ocean), N_SUB)
    _check_momentum(mesh, step, out, want, w0)


def test_quadratic_drag_steady_state_oracle():
    """Quadratic drag has no closed-form transient, but its steady state is the root of
    tau + C |w_o - w| (w_o - w) - i m f w = 0  (velocity_solver.F:3046-3056 with :3172-3203); solved here by Newton
    iteration in complex arithmetic and compared with the oracle after it has converged."""
    mesh, var, step, _, w0 = _momentum_case("hex20", True)
    nV = mesh.nVertices
    m, f, area = 917.0 * 1.9, 1.46e-4, 0.93
    tau, wo = 0.11 + 0.04j, 0.08 - 0.05j
    C = DRAGIO * RHOW * area
    out = common.run_oracle(mesh, var, step, _opts(ocean_stress_type="quadratic"), 6000)
    got = (out["uVelocity"][:nV] + 1j * out["vVelocity"][:nV])[step["solveVelocity"][:nV] == 1]
    w = got[0]
    residual = tau + C * abs(wo - w) * (wo - w) - 1j * m * f * w
    assert abs(residual) < 1e-10 * abs(tau)
    assert np.abs(got - w).max() < 1e-14


# ------------------------------------------------------------------------------------------------------------------
# (3) the reference's own grid sequence of the operator test
# ------------------------------------------------------------------------------------------------------------------
def test_operator_convergence_reference_grid_sequence():
    """testing_and_setup/testcases/square/operators_strain_stress_divergence/create_grids.py:181-213: hex grids
    82x94, 164x188, 328x376, 656x752 on the unit square (dc = 1/80 ... 1/640).  The stress-divergence error must fall
    at first to second order over the whole sequence (the guide lines of strain_stress_divergence_scaling.py:118-131)."""
    from test_oracle_kat import _operator_setup, _use_vertex, _l2
    from mpas_seaice_b200 import meshgen
    errs = []
    for nx, ny in ((82, 94), (164, 188), (328, 376), (656, 752)):
        dc = 1.0 / (nx - 2)
        mesh = meshgen.planar_hex(nx, ny, dc)
        var = oracle.init_variational(mesh, basis="wachspress", metric=False)
        ana = synthetic.operator_test_fields(mesh)
        step, opts = _operator_setup(mesh, ana["u"], ana["v"])
        oracle.subcycle_velocity_solver(mesh, var, step, opts, 1)
        nV = mesh.nVertices
        use = _use_vertex_fast(mesh) & (step["solveVelocity"][:nV] == 1)
        area = mesh.areaTriangle[:nV]
        errs.append((_l2(step["stressDivergenceU"][:nV], ana["divu"][:nV], area, use),
                     _l2(step["stressDivergenceV"][:nV], ana["divv"][:nV], area, use)))
    errs = np.array(errs)
    order = np.log2(errs[:-1] / errs[1:])
    assert np.all(order > 0.9) and np.all(order < 2.6), (errs, order)
    assert np.all(errs[-1] < 2e-3), errs


def _use_vertex_fast(mesh):
    """get_use_vertex (strain_stress_divergence_scaling.py:91-114), vectorised: drop every vertex of every cell that
    is a neighbour of a non-interior cell."""
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    coc = mesh.cellsOnCell[:nC]
    n = mesh.nEdgesOnCell[:nC]
    slot = np.arange(M)[None, :] < n[:, None]
    interior_cell = np.all(~slot | (coc <= nC), axis=1)
    nb = coc[~interior_cell][slot[~interior_cell]] - 1
    nb = np.unique(nb[nb < nC])
    use = np.ones(nV, dtype=bool)
    use[(mesh.verticesOnCell[nb][slot[nb]] - 1)] = False
    return use
