"""GPU parity of the weak operators (config_strain_scheme / config_stress_divergence_scheme = 'weak',
src/shared/mpas_seaice_velocity_solver_weak.F) and of the weak-strain + variational-divergence mix
(interpolate_strains_weak_to_variational, velocity_solver.F:2877-2972) against the oracle.  Bit-exact."""
import numpy as np
import pytest

import common
from mpas_seaice_b200 import weakmesh

pytestmark = pytest.mark.gpu

WEAK_CELL = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak",
             "replacementPressureWeak")


def _run(mesh, var, weak, step, opts, nsub):
    from mpas_seaice_b200 import host
    nC = mesh.nCells
    solver = host.EvpSolver(mesh, var, opts)
    try:
        solver.update_step(step)
        with pytest.raises(host.EvpError, match="evp_set_weak_mesh"):
            solver.run_subcycles(1)
        solver.set_weak_mesh(mesh, weak)
        solver.update_weak_state({k: np.zeros(nC + 1) for k in WEAK_CELL[:3]})
        solver.run_subcycles(nsub)
        out = solver.fetch()
        out.update(solver.fetch_weak())
        launches = solver.launch_count(nsub)
    finally:
        solver.destroy()
    return out, launches


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico4"])
@pytest.mark.parametrize("cr", ["evp", "evp_revised", "linear"])
def test_weak_weak_matches_oracle(evp_lib, kind, cr):
    mesh, var = common.mesh_case(kind)
    weak = weakmesh.weak_fields(mesh)
    step, opts = common.step_case(mesh, constitutive_relation_type=cr)
    opts = dict(opts, strain_scheme="weak", stress_divergence_scheme="weak")
    nC, nV = mesh.nCells, mesh.nVertices
    if cr == "linear":      # operator-test style: the velocity stays what the host gave
        x = np.arange(nV + 1, dtype=np.float64)
        step["uVelocity"] = np.where(step["solveVelocity"] == 1, 0.1 * np.sin(0.37 * x), 0.0)
        step["vVelocity"] = np.where(step["solveVelocity"] == 1, 0.1 * np.cos(0.11 * x), 0.0)
    nsub = 40
    ref = common.run_oracle(mesh, dict(var, weak=weak), step, opts, nsub)
    out, launches = _run(mesh, var, weak, step, opts, nsub)
    assert launches == 2 * nsub
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(out[k][vm], ref[k][vm]), k
    for k in WEAK_CELL:
        assert np.array_equal(out[k][:nC], ref[k][:nC]), k
    assert np.abs(ref["stress11Weak"]).max() > 0 and np.abs(ref["stressDivergenceU"]).max() > 0


@pytest.mark.parametrize("kind", ["hex20", "ico4"])
def test_weak_strain_variational_divergence_matches_oracle(evp_lib, kind):
    mesh, var = common.mesh_case(kind)
    weak = weakmesh.weak_fields(mesh)
    step, opts = common.step_case(mesh)
    opts = dict(opts, strain_scheme="weak", stress_divergence_scheme="variational")
    nsub = 40
    ref = common.run_oracle(mesh, dict(var, weak=weak), step, opts, nsub)
    out, launches = _run(mesh, var, weak, step, opts, nsub)
    assert launches == 5 * nsub
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(out[k][vm], ref[k][vm]), k
    for k in common.COMPARE_CELL:
        assert np.array_equal(out[k][cm], ref[k][cm]), k
    nC = mesh.nCells
    for k in ("strain11Weak", "strain22Weak", "strain12Weak"):
        assert np.array_equal(out[k][:nC], ref[k][:nC]), k


def test_invalid_scheme_combination_is_rejected(evp_lib):
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh)
    with pytest.raises(host.EvpError, match="not a valid combination"):
        host.EvpSolver(mesh, var, dict(opts, strain_scheme="variational", stress_divergence_scheme="weak"))


def test_weak_full_dynamics_step_on_device(evp_lib):
    """pre-subcycle + weak subcycle + post-subcycle on the device: init_subcycle_variables' weak branch
    (velocity_solver.F:2350-2365), seaice_final_divergence_shear_weak (weak.F:651-751, incl. its last-cell Delta) and
    the weak principal stresses (velocity_solver.F:3500-3515)."""
    import oracle
    from mpas_seaice_b200 import host, synthetic, variational_init
    mesh, var = common.mesh_case("ico4")
    weak = weakmesh.weak_fields(mesh)
    state = synthetic.sphere_state(mesh, "B")
    nC, nV = mesh.nCells, mesh.nVertices
    ref = oracle.pre_subcycle(mesh, state, 3600.0)
    _, opts = synthetic.pre_subcycle(mesh, state, 3600.0)
    opts = dict(opts, strain_scheme="weak", stress_divergence_scheme="weak")
    oracle.subcycle_velocity_solver(mesh, dict(var, weak=weak), ref, opts, 60)
    L = oracle.lib()
    want = {k: np.zeros(nC + 1) for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "p1", "p2")}
    L.orc_final_divergence_shear_weak(nC, oracle._p(ref["strain11Weak"]), oracle._p(ref["strain22Weak"]),
                                      oracle._p(ref["strain12Weak"]), oracle._p(want["divergence"]), oracle._p(want["shear"]),
                                      oracle._p(want["ridgeConvergence"]), oracle._p(want["ridgeShear"]))
    one = np.ones(nC + 1, dtype=np.int32)
    L.orc_principal_stresses_variational(nC, 1, oracle._p(one), oracle._p(ref["stress11Weak"]), oracle._p(ref["stress22Weak"]),
                                         oracle._p(ref["stress12Weak"]), oracle._p(ref["replacementPressureWeak"]),
                                         oracle._p(want["p1"]), oracle._p(want["p2"]))
    cells = synthetic.cell_inputs(state)
    cells["icePressure"] = oracle.hibler_strength_unmasked(state, nC)         # same libm exp() on both sides
    solver = host.EvpSolver(mesh, var, opts)
    try:
        solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.set_weak_mesh(mesh, weak)
        solver.pre_subcycle(cells, cold_start=True)
        solver.run_subcycles(60)
        got = solver.post_subcycle(names=("uVelocity", "vVelocity", "divergence", "shear", "ridgeConvergence", "ridgeShear",
                                          "principalStress1Weak", "principalStress2Weak"))
        wk = solver.fetch_weak()
    finally:
        solver.destroy()
    vm = ref["solveVelocity"][:nV] == 1
    assert vm.any() and not vm.all()
    for k in ("uVelocity", "vVelocity"):
        assert np.array_equal(got[k][:nV][vm], ref[k][:nV][vm]), k
    for k in ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "replacementPressureWeak"):
        assert np.array_equal(wk[k][:nC], ref[k][:nC]), k
    for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear"):
        assert np.array_equal(got[k][:nC], want[k][:nC]), k
    assert np.array_equal(got["principalStress1Weak"][:nC], want["p1"][:nC])
    assert np.array_equal(got["principalStress2Weak"][:nC], want["p2"][:nC])
    assert np.abs(want["ridgeShear"]).max() > 0


def test_pure_weak_configuration_without_variational_fields(evp_lib):
    """pkgVariational inactive in the host (config_strain_scheme = config_stress_divergence_scheme = 'weak'): no basis
    arrays, no cellVerticesAtVertex, no variational stresses cross the boundary at all."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("ico4")
    weak = weakmesh.weak_fields(mesh)
    step, opts = common.step_case(mesh)
    opts = dict(opts, strain_scheme="weak", stress_divergence_scheme="weak")
    nC, nV = mesh.nCells, mesh.nVertices
    ref = common.run_oracle(mesh, dict(var, weak=weak), step, opts, 40)
    s2 = {k: v for k, v in step.items() if k not in ("stress11", "stress22", "stress12")}
    solver = host.EvpSolver(mesh, {}, opts)
    try:
        solver.set_weak_mesh(mesh, weak)
        solver.update_step(s2)
        solver.update_weak_state({k: np.zeros(nC + 1) for k in WEAK_CELL[:3]})
        solver.run_subcycles(40)
        out = solver.fetch(names=common.COMPARE_VERTEX)
        out.update(solver.fetch_weak())
    finally:
        solver.destroy()
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(out[k][vm], ref[k][vm]), k
    for k in WEAK_CELL:
        assert np.array_equal(out[k][:nC], ref[k][:nC]), k
