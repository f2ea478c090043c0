"""CPU checks of the oracle's pre-/post-subcycle restatement (the checker of tests/test_gpu_prepost.py):
known answers of the post-subcycle routines and agreement of the two independent host restatements
(oracle C vs the product's numpy mirror) on a moving ice edge with carried state."""
import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import synthetic, variational_init


def _moving_edge_state(mesh, base, lat0):
    nC = mesh.nCells
    state = dict(base)
    cap = (np.degrees(mesh.latCell[:nC]) > lat0) | (np.degrees(mesh.latCell[:nC]) < -60.0)
    for k in ("iceAreaCell", "iceVolumeCell"):
        a = np.zeros(nC + 1)
        a[:nC] = np.where(cap, 1.0, 0.0)
        state[k] = a
    return state


def test_pre_subcycle_with_carried_state_two_restatements_agree():
    """new_ice_velocities (velocity_solver.F:1250-1279) with solveVelocityPrevious from the previous step."""
    mesh, var = common.mesh_case("ico3")
    base = synthetic.sphere_state(mesh, "B")
    prev_o = prev_n = None
    nV = mesh.nVertices
    seen_new = seen_lost = False
    for lat0 in (70.0, 55.0, 78.0):
        state = _moving_edge_state(mesh, base, lat0)
        g = oracle.pre_subcycle(mesh, state, 3600.0, prev=prev_o)
        f, opts = synthetic.pre_subcycle(mesh, state, 3600.0, prev=prev_n)
        vm = g["solveVelocity"][:nV] == 1
        for k in ("solveStress", "solveVelocity", "solveVelocityPrevious"):
            assert np.array_equal(f[k], g[k]), k
        for k in ("uVelocity", "vVelocity", "uVelocityInitial", "vVelocityInitial", "oceanStressU", "surfaceTiltForceV"):
            assert np.array_equal(f[k][:nV][vm], g[k][:nV][vm]), k
        assert np.all(g["uVelocity"][:nV][~vm] == 0.0)
        for k in ("stress11", "stress22", "stress12"):
            assert np.array_equal(f[k], g[k]), k
        if prev_o is not None:
            seen_new |= bool(((g["solveVelocity"] == 1) & (prev_o["solveVelocityPrevious"] == 0)).any())
            seen_lost |= bool(((g["solveVelocity"] == 0) & (prev_o["solveVelocityPrevious"] == 1)).any())
        oracle.subcycle_velocity_solver(mesh, var, g, opts, 5)
        prev_o = {k: g[k] for k in ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "solveVelocityPrevious")}
        prev_n = {k: g[k].copy() for k in prev_o}
    assert seen_new and seen_lost


def test_vertex_to_cell_reproduces_constants_and_skips_boundary_vertices():
    """seaice_interpolate_vertex_to_cell (mesh.F:2906-2976): weights areaTriangle * interiorVertex."""
    mesh, _ = common.mesh_case("hex20")
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    interior = variational_init.interior_vertex(mesh)
    v = np.full(nV + 1, 3.25)
    v[:nV][interior[:nV] == 0] = 1e30          # must not contaminate: weight 0
    out = np.zeros(nC + 1)
    oracle.lib().orc_interpolate_vertex_to_cell(nC, M, oracle._p(mesh.nEdgesOnCell), oracle._p(mesh.verticesOnCell),
                                                oracle._p(mesh.areaTriangle), oracle._p(interior), oracle._p(v), oracle._p(out))
    has_interior = np.array([interior[mesh.verticesOnCell[c, :mesh.nEdgesOnCell[c]] - 1].any() for c in range(nC)])
    assert np.allclose(out[:nC][has_interior], 3.25, rtol=1e-15)
    assert np.all(out[:nC][~has_interior] == 0.0)


def test_ocean_stress_final_known_answers():
    mesh, var = common.mesh_case("ico3")
    step, opts = common.step_case(mesh)
    interior = variational_init.interior_vertex(mesh)
    nV, nC = mesh.nVertices, mesh.nCells
    vm = step["solveVelocity"][:nV] == 1
    # ice moving exactly with the ocean: no stress at all
    s = common.clone_step(step)
    s["uVelocity"], s["vVelocity"] = s["uOceanVelocityVertex"].copy(), s["vOceanVelocityVertex"].copy()
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, s, opts, interior)
    assert np.all(osu[:nV][vm] == 0.0) and np.all(ocu == 0.0) and np.all(coef[:nV][vm] == 0.0)
    # ice at rest: quadratic drag  tau = c_d rho_w |u_o| u_o  per unit ice area at the vertices
    s = common.clone_step(step)
    s["uVelocity"][:] = 0.0
    s["vVelocity"][:] = 0.0
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, s, opts, interior)
    uo, vo, a = s["uOceanVelocityVertex"][:nV][vm], s["vOceanVelocityVertex"][:nV][vm], s["iceAreaVertex"][:nV][vm]
    speed = np.sqrt(uo * uo + vo * vo)
    assert np.allclose(osu[:nV][vm], 0.00536 * 1026.0 * speed * uo * a, rtol=1e-13)
    assert np.allclose(coef[:nV][vm], 0.00536 * 1026.0 * a * speed, rtol=1e-15)
    assert np.abs(ocu[:nC]).max() > 0
    # config_use_ocean_stress = false
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, s, dict(opts, use_ocean_stress=False), interior)
    assert not osu.any() and not ocu.any() and not coef.any()


def test_principal_stresses_and_divergence_shear_known_answers():
    mesh, _ = common.mesh_case("hex20")
    nC, M = mesh.nCells, mesh.maxEdges
    step = dict(solveStress=np.ones(nC + 1, dtype=np.int32),
                stress11=np.full((nC + 1, M), -3.0), stress22=np.full((nC + 1, M), -1.0), stress12=np.zeros((nC + 1, M)),
                replacementPressure=np.full((nC + 1, M), 2.0),
                strain11=np.full((nC + 1, M), 2e-7), strain22=np.full((nC + 1, M), -5e-7), strain12=np.zeros((nC + 1, M)))
    p1, p2 = oracle.principal_stresses(mesh, step)
    n = mesh.nEdgesOnCell[:nC]
    valid = np.arange(M)[None, :] < n[:, None]
    assert np.all(p1[:nC][valid] == -0.5) and np.all(p2[:nC][valid] == -1.5)
    step["replacementPressure"][:] = 0.0
    p1, _ = oracle.principal_stresses(mesh, step)
    assert np.all(p1[:nC][valid] == 1.0e30)
    ds = oracle.final_divergence_shear(mesh, step)
    div = (2e-7 - 5e-7)
    assert np.allclose(ds["divergence"][:nC], div * 100.0 * 86400.0, rtol=1e-14)
    assert np.allclose(ds["shear"][:nC], 7e-7 * 100.0 * 86400.0, rtol=1e-14)
    assert np.allclose(ds["ridgeConvergence"][:nC], -div, rtol=1e-14)
    delta = np.sqrt(div * div + (7e-7) ** 2 / 4.0)
    assert np.allclose(ds["ridgeShear"][:nC], 0.5 * (delta - abs(div)), rtol=1e-13)


def test_hibler_strength_unmasked_is_the_masked_one_where_solved():
    mesh, _ = common.mesh_case("ico3")
    state = synthetic.sphere_state(mesh, "B")
    ref = oracle.pre_subcycle(mesh, state, 3600.0)
    P = oracle.hibler_strength_unmasked(state, mesh.nCells)
    on = ref["solveStress"][:mesh.nCells] == 1
    assert np.array_equal(P[:mesh.nCells][on], ref["icePressure"][:mesh.nCells][on])
