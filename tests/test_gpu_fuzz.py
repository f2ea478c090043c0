"""Fuzz of the EVP device path against the oracle: random combinations of mesh, namelist options, masks, perturbed
forcing and subcycle counts, bit-exact like the fixed cases of test_gpu_parity.py.  (The same idea found the one
unhandled configuration of the transport kernels, tests/test_ir_parity.py::test_random_states_and_velocities_match_oracle.)
First device run: round 1's driver GPUTEST (all cases bit-identical); a difference fails the suite."""
import numpy as np
import pytest

import common
from test_gpu_parity import _compare

pytestmark = [pytest.mark.gpu]


def _random_case(seed):
    rng = np.random.default_rng(4000 + seed)
    kind = ["hex20", "quad40", "ico3", "ico4"][seed % 4]
    mesh, var = common.mesh_case(kind)
    cr = rng.choice(["evp", "evp", "evp_revised", "linear"])
    state = "auto" if kind.startswith(("hex", "quad")) else rng.choice(["A", "B"])
    step, opts = common.step_case(mesh, state_kind=state, constitutive_relation_type=str(cr),
                                  use_ocean_stress=bool(rng.uniform() < 0.8))
    opts = dict(opts, ocean_stress_type=str(rng.choice(["quadratic", "linear"])),
                average_variational_strain=bool(rng.uniform() < 0.3))
    nC, nV = mesh.nCells, mesh.nVertices
    # knock random holes into the masks (inactive cells next to active ones, isolated active vertices)
    drop_c = rng.uniform(size=nC) < rng.uniform(0.0, 0.4)
    step["solveStress"][:nC][drop_c] = 0
    drop_v = rng.uniform(size=nV) < rng.uniform(0.0, 0.4)
    step["solveVelocity"][:nV][drop_v] = 0
    # perturb the forcing and start from a non-zero state
    for k in ("airStressVertexU", "airStressVertexV", "surfaceTiltForceU", "surfaceTiltForceV", "uOceanVelocityVertex",
              "vOceanVelocityVertex", "icePressure", "totalMassVertex", "iceAreaVertex"):
        step[k] = step[k] * rng.uniform(0.5, 1.5, step[k].shape)
    on_v = step["solveVelocity"] == 1
    step["uVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
    step["vVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
    if "uVelocityInitial" in step:
        step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
    on_c = (step["solveStress"] == 1)[:, None]
    for k in ("stress11", "stress22", "stress12"):
        step[k] = np.where(on_c, rng.uniform(-500.0, 500.0, step[k].shape), 0.0)
    n_sub = int(rng.integers(1, 9))
    return mesh, var, step, opts, n_sub


def _run_device(mesh, var, step, opts, n_sub):
    from mpas_seaice_b200 import host, variational_init
    solver = host.EvpSolver(mesh, var, opts)
    try:
        if opts.get("average_variational_strain"):      # the vertex average needs areaCell (evp_set_mesh_ext)
            solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.update_step(step)
        solver.run_subcycles(n_sub)
        return solver.fetch()
    finally:
        solver.destroy()


@pytest.mark.parametrize("seed", range(16))
def test_random_configurations_match_oracle(evp_lib, seed):
    mesh, var, step, opts, n_sub = _random_case(seed)
    ref = common.run_oracle(mesh, var, step, opts, n_sub)
    out = _run_device(mesh, var, step, opts, n_sub)
    _compare(mesh, step, ref, out)
