"""Device legs of the reference-executed fixtures that were generated after this round's GPU time was spent
(tests/golden/cpu/): the oracle and the host functions replay them in tests/test_refexec_step.py / test_refexec_init.py;
here the library does, through the C ABI, bit for bit.  This file sorts last on purpose: these are the GPU tests that
have not run on a B200 yet.

* Ice shelves: two whole steps with landIceMask = 1 on a patch of the ice cover (refexec_step_*_landice_3.npz;
  init_ice_shelve_vertex_mask velocity_solver.F:481-544, the calculation masks :1023 / :1131) through
  evp_set_mesh_ext(landIceMaskVertex) and evp_pre_subcycle(landIceMask).
* Whole steps under the subcycle's other namelist options (revised EVP; linear drag with averaged strains; no ocean
  stress) and with the weak operators (pre-subcycle, weak subcycles, weak post-subcycle).
* The quadrature rules added late ('fekete', dunavant order 12: refexec_init_*.npz): evp_precompute_wachspress with
  integrationType 2 / order 12 -- the kernel is the one the other rules run, only the constant tables differ, and
  those are checked against the reference's without a GPU in tests/test_quadrature_rules.py."""
import glob
import os

import numpy as np
import pytest

from mpas_seaice_b200 import variational_init
import oracle
import test_refexec_init
import test_refexec_step
from test_refexec_step import _load, _device_step

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "cpu", "refexec_step_*landice*.npz")))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[13:-4] for f in FILES])
def test_device_reproduces_the_reference_executed_step_with_ice_shelves(evp_lib, path):
    from mpas_seaice_b200 import host
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, _ = _load(path)
    switches = dict(opts["_switches"])
    land = np.ascontiguousarray(switches.pop("land_ice_mask"), dtype=np.int32)
    land_vertex_ref = switches.pop("land_ice_mask_vertex")
    land_vertex = variational_init.land_ice_mask_vertex(mesh, land)
    assert np.array_equal(land_vertex[:mesh.nVertices], land_vertex_ref[:mesh.nVertices])
    solver = host.EvpSolver(mesh, var, {k: v for k, v in opts.items() if not k.startswith("_")})
    solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh), land_ice_mask_vertex=land_vertex)
    try:
        _device_step(solver, host, mesh, cat, dict(forcing, landIceMask=land), pre, out, nsub, host.START_FIRST_STEP, switches)
    finally:
        solver.destroy()
    lv = land_vertex[:mesh.nVertices] == 1
    assert lv.any() and not pre["solveVelocity"][:mesh.nVertices][lv].any()


OPTION_FILES = [f for f in sorted(glob.glob(os.path.join(HERE, "golden", "cpu", "refexec_step_*.npz")))
                if "landice" not in f and f not in test_refexec_step.WEAK_FILES]


@pytest.mark.gpu
@pytest.mark.parametrize("path", OPTION_FILES, ids=[os.path.basename(f)[13:-4] for f in OPTION_FILES])
def test_device_reproduces_the_reference_executed_step_with_other_options(evp_lib, path):
    """Whole steps under revised EVP and under linear drag + averaged variational strains (refexec_step_ico2_revised_4,
    refexec_step_hex12_lineardrag_avg_4): evp_aggregate -> evp_pre_subcycle -> evp_run_subcycles -> evp_post_subcycle."""
    from mpas_seaice_b200 import host
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, _ = _load(path)
    solver = host.EvpSolver(mesh, var, {k: v for k, v in opts.items() if not k.startswith("_")})
    solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
    try:
        if opts.get("strain_scheme", "variational") == "weak":
            from mpas_seaice_b200 import weakmesh
            solver.set_weak_mesh(mesh, weakmesh.weak_fields(mesh))
        _device_step(solver, host, mesh, cat, forcing, pre, out, nsub, host.START_FIRST_STEP, opts["_switches"])
    finally:
        solver.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("path", test_refexec_step.WEAK_FILES, ids=test_refexec_step.WEAK_IDS)
def test_device_reproduces_the_reference_executed_weak_step(evp_lib, path):
    """The weak operators through a whole step on the device (refexec_step_*_weak_*.npz): evp_aggregate, evp_pre_subcycle
    (weak branch of init_subcycle_variables), the weak subcycles, evp_post_subcycle (seaice_final_divergence_shear_weak,
    the weak principal stresses, ocean_stress_final) against the arrays the reference's own statements wrote."""
    from mpas_seaice_b200 import host, weakmesh
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, _ = _load(path)
    nC, nV = mesh.nCells, mesh.nVertices
    solver = host.EvpSolver(mesh, var, {k: v for k, v in opts.items() if not k.startswith("_")})
    try:
        solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.set_weak_mesh(mesh, weakmesh.weak_fields(mesh))
        solver.aggregate(cat["iceAreaCategory"][:, :, 0].copy(), cat["iceVolumeCategory"][:, :, 0].copy(),
                         cat["snowVolumeCategory"][:, :, 0].copy(), hibler_strength=True)
        agg = solver.fetch_aggregate(ice_pressure=True)
        for k in ("iceAreaCell", "iceVolumeCell", "snowVolumeCell", "totalMassCell"):
            assert np.array_equal(agg[k][:nC], pre[k][:nC]), k
        # exp() of the Hibler strength: the device's to 1 ulp of libm's, the step continues from the libm value
        p_host = oracle.hibler_strength_unmasked(dict(iceAreaCell=agg["iceAreaCell"], iceVolumeCell=agg["iceVolumeCell"]), nC)
        assert np.all(np.abs(agg["icePressure"][:nC] - p_host[:nC]) <= np.spacing(np.abs(p_host[:nC])))
        cells = dict(forcing, iceAreaCellInitial=agg["iceAreaCell"], iceAreaCell=agg["iceAreaCell"],
                     totalMassCell=agg["totalMassCell"], icePressure=p_host)
        solver.pre_subcycle(cells, cold_start=host.START_FIRST_STEP, **opts["_switches"])
        got_pre = solver.fetch_pre()
        vm = pre["solveVelocity"][:nV] == 1
        assert np.array_equal(got_pre["solveStress"][:nC], pre["solveStress"][:nC])
        assert np.array_equal(got_pre["solveVelocity"][:nV], pre["solveVelocity"][:nV])
        for k in ("totalMassVertexfVertex", "airStressVertexU", "airStressVertexV", "surfaceTiltForceU", "surfaceTiltForceV",
                  "oceanStressU", "oceanStressV", "uVelocityInitial", "vVelocityInitial"):
            assert np.array_equal(got_pre[k][:nV][vm], pre[k][:nV][vm]), k
        solver.run_subcycles(nsub)
        got = solver.post_subcycle(names=("uVelocity", "vVelocity", "divergence", "shear", "ridgeConvergence", "ridgeShear",
                                          "oceanStressCellU", "oceanStressCellV", "oceanStressU", "oceanStressV",
                                          "oceanStressCoeff", "principalStress1Weak", "principalStress2Weak"))
        wk = solver.fetch_weak()
    finally:
        solver.destroy()
    for k in test_refexec_step.WEAK_CELL:
        assert np.array_equal(wk[k][:nC], out[k][:nC]), k
    for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV",
              "principalStress1Weak", "principalStress2Weak"):
        assert np.array_equal(got[k][:nC], out[k][:nC]), k
    for k in ("uVelocity", "vVelocity", "oceanStressU", "oceanStressV", "oceanStressCoeff"):
        assert np.array_equal(got[k][:nV][vm], out[k][:nV][vm]), k
    assert np.abs(out["ridgeShear"][:nC]).max() > 0 and np.abs(out["stress12Weak"][:nC]).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", test_refexec_init.CPU_FILES, ids=[os.path.basename(f)[13:-4] for f in test_refexec_init.CPU_FILES])
def test_device_precompute_reproduces_the_reference_executed_arrays_of_the_late_rules(evp_lib, path):
    import common
    from mpas_seaice_b200 import host
    mesh, want, kw, _ = test_refexec_init._load(path)
    var = oracle.init_variational(mesh, **kw)            # xLocal / yLocal and the host-side maps the create call takes
    step, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]),
                            integration=(kw["integration_type"], kw["integration_order"]))
    try:
        got = solver.fetch_basis()
    finally:
        solver.destroy()
    nC = mesh.nCells
    for k, a in got.items():
        assert np.array_equal(a[:nC], want[k][:nC]), k


# ---------------------------------------------------------------------------------------------------------------------
# Fuzz of what a handle keeps BETWEEN calls (tile and vertex-block lists, contrib rows, the instantiated graph or the
# persistent kernel, the resident state): sequences of steps with a different random ice cover each.  200 seeds of each
# kind ran on the emulated library when these were written (persistent and graph path): all bit-identical.
# ---------------------------------------------------------------------------------------------------------------------
def _random_cover(rng, mesh):
    nC = mesh.nCells
    mode = rng.integers(0, 5)
    if mode == 0:                                   # random holes
        on = rng.uniform(size=nC) < rng.uniform(0.05, 0.95)
    elif mode == 1:                                 # an ice edge
        x = mesh.latCell[:nC] if mesh.on_a_sphere else mesh.xCell[:nC] / mesh.xCell[:nC].max()
        t = rng.uniform(x.min(), x.max())
        on = x > t if rng.uniform() < 0.5 else x < t
    elif mode == 2:                                 # a run of cell indices: whole tiles without work
        lo = int(rng.integers(0, nC))
        on = np.zeros(nC, bool)
        on[lo:min(nC, lo + int(rng.integers(1, nC)))] = True
    elif mode == 3:
        on = np.ones(nC, bool)
    else:                                           # a few isolated cells
        on = np.zeros(nC, bool)
        on[rng.integers(0, nC, size=int(rng.integers(1, 6)))] = True
    area = np.where(on, rng.uniform(0.2, 1.0, nC), 0.0)
    return area, np.where(on, area * rng.uniform(0.2, 3.0, nC), 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_random_ice_cover_sequences_with_the_state_resident(evp_lib, seed):
    """Five whole steps (evp_pre_subcycle -> evp_run_subcycles -> evp_post_subcycle) on one handle, the ice cover redrawn
    every step, u / v / stresses / solveVelocityPrevious never re-uploaded; the oracle chain carries them explicitly.
    Compared on ALL vertices and cells (what new_ice_velocities zeroes, what a cell that lost its ice keeps)."""
    import test_gpu_prepost as P
    from mpas_seaice_b200 import host, synthetic
    rng = np.random.default_rng(7000 + seed)
    mesh, var = common_mesh(["hex20", "ico3", "quad40", "ico4"][seed % 4])
    base = P._state(mesh, "B" if mesh.on_a_sphere else "square")
    nC, nV = mesh.nCells, mesh.nVertices
    interior = variational_init.interior_vertex(mesh)
    _, opts = synthetic.pre_subcycle(mesh, base, 3600.0, constitutive_relation_type=str(rng.choice(["evp", "evp_revised"])))
    opts = dict(opts, ocean_stress_type=str(rng.choice(["quadratic", "linear"])))
    valid = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
    solver = P._solver(mesh, var, opts)
    prev = None
    try:
        for it in range(5):
            state = dict(base)
            area, vol = _random_cover(rng, mesh)
            for k, a in (("iceAreaCell", area), ("iceVolumeCell", vol), ("snowVolumeCell", 0.1 * vol)):
                z = np.zeros(nC + 1)
                z[:nC] = a
                state[k] = z
            n_sub = int(rng.integers(1, 8))
            ref_step = oracle.pre_subcycle(mesh, state, 3600.0, prev=prev)
            solver.pre_subcycle(P._cells(mesh, state), cold_start=(it == 0))
            got_pre = solver.fetch_pre()
            for k, n in (("solveStress", nC), ("solveVelocity", nV), ("solveVelocityPrevious", nV)):
                assert np.array_equal(got_pre[k][:n], ref_step[k][:n]), (it, k)
            oracle.subcycle_velocity_solver(mesh, var, ref_step, opts, n_sub)
            solver.run_subcycles(n_sub)
            ref = P._post_reference(mesh, ref_step, opts, interior)
            got = solver.post_subcycle(names=host.POST_FIELDS_VARIATIONAL)
            inner = solver.fetch(names=("stress11", "stress22", "stress12"))
            vm = ref_step["solveVelocity"][:nV] == 1
            for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV"):
                assert np.array_equal(got[k][:nC], ref[k][:nC]), (it, k)
            for k in ("uVelocity", "vVelocity"):
                assert np.array_equal(got[k][:nV], ref[k][:nV]), (it, k)
            for k in ("oceanStressU", "oceanStressV", "oceanStressCoeff"):
                assert np.array_equal(got[k][:nV][vm], ref[k][:nV][vm]), (it, k)
            for k in ("stress11", "stress22", "stress12"):
                assert np.array_equal(inner[k][:nC][valid], ref_step[k][:nC][valid]), (it, k)
            prev = {k: ref_step[k] for k in ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "solveVelocityPrevious")}
    finally:
        solver.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_random_mask_sequences_on_one_handle(evp_lib, seed):
    """evp_update_step -> evp_run_subcycles (sometimes split in two calls) -> evp_fetch, five times on one handle with new
    masks, state and subcycle count each time: nothing a previous step left on the device may show."""
    from mpas_seaice_b200 import host
    from test_gpu_parity import _compare
    import common
    rng = np.random.default_rng(9000 + seed)
    mesh, var = common_mesh(["hex20", "quad40", "ico3", "ico4"][seed % 4])
    state_kind = "auto" if not mesh.on_a_sphere else str(rng.choice(["A", "B"]))
    base, opts = common.step_case(mesh, state_kind=state_kind, constitutive_relation_type=str(rng.choice(["evp", "evp_revised"])))
    opts = dict(opts, ocean_stress_type=str(rng.choice(["quadratic", "linear"])),
                average_variational_strain=bool(rng.uniform() < 0.2))
    nC, nV = mesh.nCells, mesh.nVertices
    solver = host.EvpSolver(mesh, var, opts)
    try:
        if opts["average_variational_strain"]:
            solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        for it in range(5):
            step = common.clone_step(base)
            mode = rng.integers(0, 4)
            if mode == 0:
                step["solveStress"][:nC][rng.uniform(size=nC) < rng.uniform(0, 0.9)] = 0
                step["solveVelocity"][:nV][rng.uniform(size=nV) < rng.uniform(0, 0.9)] = 0
            elif mode == 1:
                lo = int(rng.integers(0, nC))
                keep = np.zeros(nC, bool)
                keep[lo:min(nC, lo + int(rng.integers(1, nC)))] = True
                step["solveStress"][:nC][~keep] = 0
                lo = int(rng.integers(0, nV))
                keep = np.zeros(nV, bool)
                keep[lo:min(nV, lo + int(rng.integers(1, nV)))] = True
                step["solveVelocity"][:nV][~keep] = 0
            elif mode == 2:
                if rng.uniform() < 0.3:
                    step["solveStress"][:] = 0
                if rng.uniform() < 0.3:
                    step["solveVelocity"][:] = 0
            on_v = step["solveVelocity"] == 1
            step["uVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            step["vVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            if "uVelocityInitial" in step:
                step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
            on_c = (step["solveStress"] == 1)[:, None]
            for k in ("stress11", "stress22", "stress12"):
                step[k] = np.where(on_c, rng.uniform(-500.0, 500.0, step[k].shape), 0.0)
            n_sub = int(rng.integers(1, 7))
            ref = common.run_oracle(mesh, var, step, opts, n_sub)
            solver.update_step(step)
            first = int(rng.integers(0, n_sub + 1)) if rng.uniform() < 0.5 else n_sub
            if first:
                solver.run_subcycles(first)
            if n_sub - first:
                solver.run_subcycles(n_sub - first)
            _compare(mesh, step, ref, solver.fetch())
    finally:
        solver.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_options_switched_on_one_handle(evp_lib, seed):
    """evp_set_options between the steps of one handle: constitutive relation (evp / evp_revised / linear / none), drag law,
    ocean stress on / off, averaged strains and the operator schemes (variational, weak, weak strain + variational
    divergence; the weak mesh is set once) redrawn every step, masks and state too.  104 seeds x 6 steps on the emulated
    library when this was written: all bit-identical."""
    import common
    from mpas_seaice_b200 import host, weakmesh
    from test_gpu_parity import _compare
    rng = np.random.default_rng(11000 + seed)
    mesh, var = common_mesh(["hex20", "quad40", "ico3", "ico4"][seed % 4])
    weak = weakmesh.weak_fields(mesh)
    state = "auto" if not mesh.on_a_sphere else str(rng.choice(["A", "B"]))
    base, opts0 = common.step_case(mesh, state_kind=state)
    nC, nV = mesh.nCells, mesh.nVertices
    solver = host.EvpSolver(mesh, var, opts0)
    try:
        solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.set_weak_mesh(mesh, weak)
        for it in range(6):
            cr = str(rng.choice(["evp", "evp_revised", "linear", "none"]))
            _, o = common.step_case(mesh, state_kind=state, constitutive_relation_type=cr)
            scheme = [("variational", "variational"), ("weak", "weak"), ("weak", "variational")][int(rng.integers(0, 3))]
            opts = dict(o, ocean_stress_type=str(rng.choice(["quadratic", "linear"])), use_ocean_stress=bool(rng.uniform() < 0.8),
                        average_variational_strain=bool(scheme[0] == "variational" and rng.uniform() < 0.3),
                        strain_scheme=scheme[0], stress_divergence_scheme=scheme[1])
            step = common.clone_step(base)
            step["solveStress"][:nC][rng.uniform(size=nC) < rng.uniform(0, 0.5)] = 0
            step["solveVelocity"][:nV][rng.uniform(size=nV) < rng.uniform(0, 0.5)] = 0
            on_v = step["solveVelocity"] == 1
            step["uVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            step["vVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
            on_c = step["solveStress"] == 1
            for k in ("stress11", "stress22", "stress12"):
                step[k] = np.where(on_c[:, None], rng.uniform(-500.0, 500.0, step[k].shape), 0.0)
            if scheme[0] == "weak":
                for k in ("stress11Weak", "stress22Weak", "stress12Weak"):
                    step[k] = np.where(on_c, rng.uniform(-500.0, 500.0, nC + 1), 0.0)
            n_sub = int(rng.integers(1, 6))
            ref = common.run_oracle(mesh, dict(var, weak=weak), step, opts, n_sub)
            solver.set_options(opts)
            solver.update_step(step)
            if scheme[0] == "weak":
                solver.update_weak_state({k: step[k] for k in ("stress11Weak", "stress22Weak", "stress12Weak")})
            solver.run_subcycles(n_sub)
            out = solver.fetch()
            if scheme == ("weak", "weak"):
                wk = solver.fetch_weak()
                _, vm = common.masks_for(mesh, step)
                for k in ("uVelocity", "vVelocity"):
                    assert np.array_equal(out[k][vm], ref[k][vm]), (it, cr, k)
                for k in ("stress11Weak", "stress22Weak", "stress12Weak"):
                    assert np.array_equal(wk[k][:nC], ref[k][:nC]), (it, cr, k)
            else:
                _compare(mesh, step, ref, out)
    finally:
        solver.destroy()


_RULES = [("dunavant", o) for o in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)] + [("fekete", o) for o in (1, 2, 3, 4, 5, 6, 8, 9)] + \
    [("trapezoidal", o) for o in (1, 2, 3, 5)]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_precompute_on_distorted_cells(evp_lib, seed):
    """evp_precompute_wachspress (a random quadrature rule of the 23) / evp_precompute_pwl on cells whose vertices are
    jittered by up to 22 % of the cell spacing -- planar hexagons, quadrilaterals, the sphere -- against the oracle, bit for bit.  400 seeds on the emulated library when this was written: identical."""
    from mpas_seaice_b200 import host, meshgen
    import common
    rng = np.random.default_rng(13000 + seed)
    mesh = meshgen.Mesh([meshgen.planar_hex(8, 9, 16000.0), meshgen.planar_quad(7, 7, 16000.0), meshgen.icosphere(2)][seed % 3])
    nV = mesh.nVertices
    amp = rng.uniform(0.05, 0.22)
    if mesh.on_a_sphere:
        p = np.stack([mesh.xVertex[:nV], mesh.yVertex[:nV], mesh.zVertex[:nV]], 1)
        p = p + amp * float(mesh.dcEdge[:-1].mean()) * rng.uniform(-1, 1, (nV, 3))
        p *= (mesh.sphere_radius / np.linalg.norm(p, axis=1))[:, None]
        for k, col in (("xVertex", 0), ("yVertex", 1), ("zVertex", 2)):
            a = mesh[k].copy()
            a[:nV] = p[:, col]
            setattr(mesh, k, a)
    else:
        for k in ("xVertex", "yVertex"):
            a = mesh[k].copy()
            a[:nV] += amp * 16000.0 * rng.uniform(-1, 1, nV)
            setattr(mesh, k, a)
    basis = "pwl" if seed == 4 else "wachspress"
    itype, order = _RULES[int(rng.integers(0, len(_RULES)))]
    var = oracle.init_variational(mesh, **(dict(basis="pwl") if basis == "pwl" else dict(integration_type=itype, integration_order=order)))
    _, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]), integration=(itype, order), basis=basis)
    try:
        got = solver.fetch_basis()
    finally:
        solver.destroy()
    nC = mesh.nCells
    assert np.abs(got["basisGradientU"][:nC]).max() > 0
    for k, a in got.items():
        assert np.isfinite(var[k][:nC]).all(), k          # (a degenerate cell would give NaN on both sides, with different signs)
        assert np.array_equal(a[:nC], var[k][:nC]), k


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_random_switches_ice_shelves_and_categories(evp_lib, seed):
    """Three resident steps with the pre-subcycle's namelist switches (air stress, surface tilt, geostrophic or
    sea-surface-height tilt), an ice-shelf mask and 1-3 thickness categories (evp_aggregate) drawn at random, the cover
    redrawn every step.  132 seeds on the emulated library when this was written: all bit-identical."""
    import test_gpu_prepost as P
    from mpas_seaice_b200 import host, synthetic
    rng = np.random.default_rng(61000 + seed)
    mesh, var = common_mesh(["hex20", "ico3", "quad40", "ico4"][seed % 4])
    base = P._state(mesh, "B" if mesh.on_a_sphere else "square")
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    interior = variational_init.interior_vertex(mesh)
    _, opts = synthetic.pre_subcycle(mesh, base, 3600.0)
    sw = dict(use_air_stress=bool(rng.uniform() < 0.7), use_surface_tilt=bool(rng.uniform() < 0.7),
              geostrophic_surface_tilt=bool(rng.uniform() < 0.6))
    land = land_vertex = None
    if seed % 2 == 0:
        land = np.zeros(nC + 1, np.int32)
        land[:nC] = rng.uniform(size=nC) < rng.uniform(0.05, 0.4)
        land_vertex = variational_init.land_ice_mask_vertex(mesh, land)
    n_cat = int(rng.integers(1, 4))
    solver = host.EvpSolver(mesh, var, opts)
    solver.set_mesh_ext(mesh, interior, **({"land_ice_mask_vertex": land_vertex} if land is not None else {}))
    prev = dict(uVelocity=np.zeros(nV + 1), vVelocity=np.zeros(nV + 1), solveVelocityPrevious=np.zeros(nV + 1, dtype=np.int32),
                stress11=np.zeros((nC + 1, M)), stress22=np.zeros((nC + 1, M)), stress12=np.zeros((nC + 1, M)))
    try:
        for it in range(3):
            w = rng.uniform(0.1, 1.0, n_cat)
            w /= w.sum()
            area = np.where(rng.uniform(size=nC) < rng.uniform(0.2, 1.0), rng.uniform(0.2, 1.0, nC), 0.0)
            a, vi, vs = (np.zeros((nC + 1, n_cat)) for _ in range(3))
            for k in range(n_cat):
                a[:nC, k] = area * w[k]
                vi[:nC, k] = area * w[k] * rng.uniform(0.3, 3.0, nC)
                vs[:nC, k] = 0.1 * vi[:nC, k]
            A, VI, VS, mass = oracle.aggregate_mass_and_area(a, vi, vs)
            state = dict(base, iceAreaCell=A, iceVolumeCell=VI, snowVolumeCell=VS)
            forcing = {}
            if not sw["geostrophic_surface_tilt"]:
                forcing = dict(seaSurfaceTiltU=1e-6 * rng.uniform(-1, 1, nC + 1), seaSurfaceTiltV=1e-6 * rng.uniform(-1, 1, nC + 1))
                state.update(forcing)
            kw = dict(sw, **(dict(land_ice_mask=land, land_ice_mask_vertex=land_vertex) if land is not None else {}))
            ref_step = oracle.pre_subcycle(mesh, state, 3600.0, prev=prev, **kw)
            solver.aggregate(a.copy(), vi.copy(), vs.copy(), hibler_strength=False)
            agg = solver.fetch_aggregate(ice_pressure=False)
            for k, want in (("iceAreaCell", A), ("iceVolumeCell", VI), ("snowVolumeCell", VS), ("totalMassCell", mass)):
                assert np.array_equal(agg[k][:nC], want[:nC]), (it, k)
            cells = dict({k: np.ascontiguousarray(state[k], dtype=np.float64)
                          for k in ("uOceanVelocity", "vOceanVelocity", "uAirVelocity", "vAirVelocity", "airDensity")},
                         iceAreaCellInitial=agg["iceAreaCell"], iceAreaCell=agg["iceAreaCell"], totalMassCell=agg["totalMassCell"],
                         icePressure=oracle.hibler_strength_unmasked(state, nC), **forcing)
            if land is not None:
                cells["landIceMask"] = land
            n_sub = int(rng.integers(1, 6))
            solver.pre_subcycle(cells, cold_start=(host.START_FIRST_STEP if it == 0 else host.START_RESIDENT), **sw)
            got_pre = solver.fetch_pre()
            for k, n in (("solveStress", nC), ("solveVelocity", nV), ("solveVelocityPrevious", nV)):
                assert np.array_equal(got_pre[k][:n], ref_step[k][:n]), (it, k)
            vm = ref_step["solveVelocity"][:nV] == 1
            for k in ("airStressVertexU", "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "totalMassVertexfVertex"):
                assert np.array_equal(got_pre[k][:nV][vm], ref_step[k][:nV][vm]), (it, k)
            oracle.subcycle_velocity_solver(mesh, var, ref_step, opts, n_sub)
            solver.run_subcycles(n_sub)
            ref = P._post_reference(mesh, ref_step, opts, interior)
            got = solver.post_subcycle(names=host.POST_FIELDS_VARIATIONAL)
            for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV"):
                assert np.array_equal(got[k][:nC], ref[k][:nC]), (it, k)
            for k in ("uVelocity", "vVelocity"):
                assert np.array_equal(got[k][:nV], ref[k][:nV]), (it, k)
            prev = {k: ref_step[k] for k in ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "solveVelocityPrevious")}
    finally:
        solver.destroy()


def common_mesh(kind):
    import common
    return common.mesh_case(kind)
