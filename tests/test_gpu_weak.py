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
