"""A whole dynamics step -- the body of seaice_run_velocity_solver (src/shared/mpas_seaice_velocity_solver.F:562-595):
velocity_solver_pre_subcycle, subcycle_velocity_solver, velocity_solver_post_subcycle -- against golden vectors produced
by EXECUTING THE REFERENCE'S FORTRAN SOURCE (tests/golden/step/refexec_step_*.npz; generator
tests/golden/make_reference_executed_golden.py, interpreter tests/golden/fortran_subset.py).

Input: category tracers, coupler fields, the first step of a run (solveVelocityPrevious = 0, velocities and stresses at
rest).  Checked after the pre-subcycle and after the post-subcycle, bit for bit:
  CPU  the oracle's restatements (oracle.aggregate_mass_and_area, pre_subcycle, subcycle_velocity_solver,
       final_divergence_shear, principal_stresses, ocean_stress_final);
  GPU  the library through the C ABI (evp_aggregate for several categories, evp_pre_subcycle, evp_run_subcycles,
       evp_post_subcycle).
"""
import glob
import os

import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import synthetic, variational_init

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "step", "refexec_step_*.npz")))
IDS = [os.path.basename(f)[13:-4] for f in FILES]
# generated after this round's GPU time was spent: replayed on the oracle and the host functions only
CPU_FILES = sorted(glob.glob(os.path.join(HERE, "golden", "cpu", "refexec_step_*.npz")))
CPU_IDS = [os.path.basename(f)[13:-4] for f in CPU_FILES]

PRE_CELL = ("solveStress", "icePressure", "totalMassCell", "iceAreaCellInitial", "iceAreaCell", "iceVolumeCell", "snowVolumeCell")
PRE_VERTEX_ALL = ("solveVelocity", "iceAreaVertex", "totalMassVertex", "uOceanVelocityVertex", "vOceanVelocityVertex")
PRE_VERTEX_SOLVED = ("totalMassVertexfVertex", "airStressVertexU", "airStressVertexV", "surfaceTiltForceU", "surfaceTiltForceV",
                     "oceanStressU", "oceanStressV", "uVelocityInitial", "vVelocityInitial", "uVelocity", "vVelocity")
POST_CELL = ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV")
POST_VERTEX = ("uVelocity", "vVelocity", "oceanStressU", "oceanStressV", "oceanStressCoeff", "stressDivergenceU", "stressDivergenceV")
POST_CELL2D = ("stress11", "stress22", "stress12", "strain11", "strain22", "strain12", "replacementPressure")


def _load(path):
    z = np.load(path)
    mesh, var = common.mesh_case(str(z["kind"]))
    opts = {k[4:]: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files if k.startswith("opt_")}
    cat = {k: z["in_" + k] for k in ("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory")}
    forcing = {k: z["in_" + k] for k in ("uAirVelocity", "vAirVelocity", "airDensity", "uOceanVelocity", "vOceanVelocity")}
    sw = {k[3:]: bool(z[k]) for k in z.files if k.startswith("sw_")}          # the pre-subcycle's namelist switches
    if sw and not sw.get("geostrophic_surface_tilt", True):
        forcing.update(seaSurfaceTiltU=z["in_seaSurfaceTiltU"], seaSurfaceTiltV=z["in_seaSurfaceTiltV"])
    opts["_switches"] = sw or dict(use_air_stress=True, use_surface_tilt=True, geostrophic_surface_tilt=True)
    if "in_landIceMask" in z.files:     # ice shelves: the vertex mask is the reference's own (init_ice_shelve_vertex_mask)
        opts["_switches"].update(land_ice_mask=z["in_landIceMask"], land_ice_mask_vertex=z["ref_landIceMaskVertex"])
    pre = {k[4:]: z[k] for k in z.files if k.startswith("pre_")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    n_steps = int(z["n_steps"]) if "n_steps" in z.files else 1
    if n_steps > 1:        # further steps of the same run: (category tracers, fields after pre, fields after post) each
        more = []
        for n in range(2, n_steps + 1):
            tag = str(n)
            more.append(({k: z["in%s_%s" % (tag, k)] for k in cat},
                         {k[len("pre" + tag) + 1:]: z[k] for k in z.files if k.startswith("pre" + tag + "_")},
                         {k[len("out" + tag) + 1:]: z[k] for k in z.files if k.startswith("out" + tag + "_")}))
        pre["_more"] = more
    return mesh, var, opts, cat, forcing, pre, out, int(z["nsub"]), float(z["config_dt"]), str(z["provenance"])


def _first_step_prev(mesh):
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    return dict(uVelocity=np.zeros(nV + 1), vVelocity=np.zeros(nV + 1), solveVelocityPrevious=np.zeros(nV + 1, dtype=np.int32),
                stress11=np.zeros((nC + 1, M)), stress22=np.zeros((nC + 1, M)), stress12=np.zeros((nC + 1, M)))


def _valid(mesh):
    return np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:mesh.nCells, None]


def test_fixtures_exist_and_name_the_routines_that_ran():
    assert len(FILES) >= 3
    seen = set()
    for f in FILES:
        prov = _load(f)[-1]
        assert "interpreting the reference's Fortran source" in prov
        seen |= {w.strip() for w in prov.split(":", 1)[1].split(",")}
    for name in ("velocity_solver_pre_subcycle", "aggregate_mass_and_area", "stress_calculation_mask", "velocity_calculation_mask",
                 "new_ice_velocities", "ice_strength", "constant_air_stress", "coriolis_force_coefficient", "ocean_stress",
                 "surface_tilt_geostrophic", "init_subcycle_variables", "seaice_interpolate_cell_to_vertex",
                 "subcycle_velocity_solver", "velocity_solver_post_subcycle", "seaice_final_divergence_shear_variational",
                 "principal_stresses", "ocean_stress_final", "seaice_interpolate_vertex_to_cell"):
        assert name in seen, name


WEAK_FILES = [f for f in CPU_FILES if "_weak_" in os.path.basename(f)]
WEAK_IDS = [os.path.basename(f)[13:-4] for f in WEAK_FILES]
WEAK_CELL = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak", "replacementPressureWeak")


@pytest.mark.parametrize("path", [f for f in FILES + CPU_FILES if f not in WEAK_FILES],
                         ids=[i for i in IDS + CPU_IDS if i not in WEAK_IDS])
def test_oracle_reproduces_the_reference_executed_step(path):
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, _ = _load(path)
    prev = _first_step_prev(mesh)
    for cat_n, pre_n, out_n in [(cat, pre, out)] + pre.get("_more", []):
        prev = _oracle_step(mesh, var, opts, cat_n, forcing, pre_n, out_n, nsub, config_dt, prev)


@pytest.mark.parametrize("path", [f for f in CPU_FILES if "landice" in f], ids=[i for i in CPU_IDS if "landice" in i])
def test_ice_shelf_masks_reproduce_the_reference_executed_arrays(path):
    """init_ice_shelve_vertex_mask (velocity_solver.F:481-544) and dynamically_locked_cell_mask (:402-467), integer maps
    SURVEY section 8(c) lists for bit-exact parity: the oracle's restatements and the host functions against the arrays
    the reference's statements wrote; and the masks of the step show the shelf (no stress point, no velocity point on
    it, although there is ice)."""
    z = np.load(path)
    mesh, _ = common.mesh_case(str(z["kind"]))
    nC, nV = mesh.nCells, mesh.nVertices
    land = z["in_landIceMask"]
    interior = z["ref_interiorVertex"]
    assert np.array_equal(variational_init.interior_vertex(mesh), interior)
    for f in (oracle.land_ice_mask_vertex, variational_init.land_ice_mask_vertex):
        assert np.array_equal(f(mesh, land)[:nV], z["ref_landIceMaskVertex"][:nV]), f.__module__
    for f in (oracle.dynamically_locked_cells_mask, variational_init.dynamically_locked_cells_mask):
        assert np.array_equal(f(mesh, interior)[:nC], z["ref_dynamicallyLockedCellsMask"][:nC]), f.__module__
    locked = z["ref_dynamicallyLockedCellsMask"][:nC]
    assert not locked.any()       # every cell of these meshes touches an interior vertex (a mixed case: test_host_numpy.py)
    lv = z["ref_landIceMaskVertex"][:nV] == 1
    assert 0 < lv.sum() < nV
    assert not z["pre_solveVelocity"][:nV][lv].any()
    # a shelf cell with shelf cells all around has no stress point although it carries ice
    shelf = land[:nC] == 1
    ring = np.zeros(nC, dtype=bool)
    for k in range(mesh.maxEdges):
        nb = mesh.cellsOnCell[:nC, k] - 1
        ok = (k < mesh.nEdgesOnCell[:nC]) & (nb < nC)
        ring |= ok & ~shelf[np.minimum(nb, nC - 1)]
    deep = shelf & ~ring
    assert deep.any() and (z["pre_iceAreaCell"][:nC][deep] > 0.1).any()
    assert not z["pre_solveStress"][:nC][deep].any()
    # the host mirror of the pre-subcycle gives the reference's masks with the shelf
    state = dict(iceAreaCell=z["pre_iceAreaCell"], iceVolumeCell=z["pre_iceVolumeCell"], snowVolumeCell=z["pre_snowVolumeCell"],
                 **{k: z["in_" + k] for k in ("uAirVelocity", "vAirVelocity", "airDensity", "uOceanVelocity", "vOceanVelocity")})
    step, _ = synthetic.pre_subcycle(mesh, state, float(z["config_dt"]), n_elastic=int(z["nsub"]), land_ice_mask=land)
    assert np.array_equal(step["solveStress"][:nC], z["pre_solveStress"][:nC])
    assert np.array_equal(step["solveVelocity"][:nV], z["pre_solveVelocity"][:nV])


def _oracle_step(mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, prev):
    nC, nV = mesh.nCells, mesh.nVertices
    a, vi, vs, mass = oracle.aggregate_mass_and_area(cat["iceAreaCategory"][:, :, 0], cat["iceVolumeCategory"][:, :, 0],
                                                     cat["snowVolumeCategory"][:, :, 0])
    for k, got in (("iceAreaCell", a), ("iceVolumeCell", vi), ("snowVolumeCell", vs), ("totalMassCell", mass)):
        assert np.array_equal(got[:nC], pre[k][:nC]), k
    state = dict(forcing, iceAreaCell=a, iceVolumeCell=vi, snowVolumeCell=vs)
    step = oracle.pre_subcycle(mesh, state, config_dt, prev=prev, use_ocean_stress=bool(opts.get("use_ocean_stress", True)),
                               **opts["_switches"])
    vm = pre["solveVelocity"][:nV] == 1
    assert vm.any() and (pre["solveStress"][:nC] == 1).any()
    for k in ("solveStress", "icePressure"):
        assert np.array_equal(step[k][:nC], pre[k][:nC]), k
    for k in PRE_VERTEX_ALL + ("solveVelocityPrevious",):
        assert np.array_equal(step[k][:nV], pre[k][:nV]), k
    for k in PRE_VERTEX_SOLVED:
        assert np.array_equal(step[k][:nV][vm], pre[k][:nV][vm]), k
    if opts.get("strain_scheme", "variational") == "weak":     # weak strains, variational stress divergence
        from mpas_seaice_b200 import weakmesh
        var = dict(var, weak=weakmesh.weak_fields(mesh))
    oracle.subcycle_velocity_solver(mesh, var, step, opts, nsub)
    interior = variational_init.interior_vertex(mesh)
    ds = oracle.final_divergence_shear(mesh, step)
    p1, p2 = oracle.principal_stresses(mesh, step)
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, step, opts, interior)
    got = dict(step, **ds)
    got.update(oceanStressU=osu, oceanStressV=osv, oceanStressCellU=ocu, oceanStressCellV=ocv, oceanStressCoeff=coef,
               principalStress1=p1, principalStress2=p2)
    cm = (pre["solveStress"][:nC] == 1)[:, None] & _valid(mesh)
    for k in POST_CELL:
        assert np.array_equal(got[k][:nC], out[k][:nC]), k
    for k in POST_CELL2D + ("principalStress1", "principalStress2"):
        assert np.array_equal(got[k][:nC][cm], out[k][:nC][cm]), k
    for k in POST_VERTEX:
        assert np.array_equal(got[k][:nV][vm], out[k][:nV][vm]), k
    assert np.abs(out["uVelocity"][:nV][vm]).max() > 0 and np.abs(out["divergence"][:nC]).max() > 0
    assert np.abs(out["stress11"][:nC][cm]).max() > 0
    assert (np.abs(out["oceanStressCellU"][:nC]).max() > 0) == bool(opts.get("use_ocean_stress", True))
    # vertices that lost their ice are zeroed by the reference (new_ice_velocities :1262-1270): compare everywhere
    assert np.array_equal(step["uVelocity"][:nV][~vm], out["uVelocity"][:nV][~vm])
    return dict(uVelocity=step["uVelocity"], vVelocity=step["vVelocity"], stress11=step["stress11"], stress22=step["stress22"],
                stress12=step["stress12"], solveVelocityPrevious=step["solveVelocityPrevious"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_device_reproduces_the_reference_executed_step(evp_lib, path):
    from mpas_seaice_b200 import host
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, _ = _load(path)
    nC, nV = mesh.nCells, mesh.nVertices
    solver = host.EvpSolver(mesh, var, {k: v for k, v in opts.items() if not k.startswith("_")})
    solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
    try:
        steps = [(cat, pre, out)] + pre.get("_more", [])
        for n_step, (cat, pre, out) in enumerate(steps):
            start = host.START_FIRST_STEP if n_step == 0 else host.START_RESIDENT       # the state stays on the device
            _device_step(solver, host, mesh, cat, forcing, pre, out, nsub, start, opts["_switches"])
    finally:
        solver.destroy()


def _oracle_pre(mesh, opts, cat, forcing, pre, config_dt, prev):
    nC, nV = mesh.nCells, mesh.nVertices
    a, vi, vs, mass = oracle.aggregate_mass_and_area(cat["iceAreaCategory"][:, :, 0], cat["iceVolumeCategory"][:, :, 0],
                                                     cat["snowVolumeCategory"][:, :, 0])
    state = dict(forcing, iceAreaCell=a, iceVolumeCell=vi, snowVolumeCell=vs)
    step = oracle.pre_subcycle(mesh, state, config_dt, prev=prev, use_ocean_stress=bool(opts.get("use_ocean_stress", True)),
                               **opts["_switches"])
    vm = pre["solveVelocity"][:nV] == 1
    assert vm.any() and (pre["solveStress"][:nC] == 1).any()
    for k in ("solveStress", "icePressure"):
        assert np.array_equal(step[k][:nC], pre[k][:nC]), k
    for k in PRE_VERTEX_ALL + ("solveVelocityPrevious",):
        assert np.array_equal(step[k][:nV], pre[k][:nV]), k
    for k in PRE_VERTEX_SOLVED:
        assert np.array_equal(step[k][:nV][vm], pre[k][:nV][vm]), k
    return step, vm


def _weak_post_oracle(mesh, step):
    """seaice_final_divergence_shear_weak (weak.F:651-751) and the weak branch of principal_stresses
    (velocity_solver.F:3500-3515) from the oracle's weak state"""
    nC = mesh.nCells
    L = oracle.lib()
    want = {k: np.zeros(nC + 1) for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "principalStress1Weak",
                                          "principalStress2Weak")}
    L.orc_final_divergence_shear_weak(nC, oracle._p(step["strain11Weak"]), oracle._p(step["strain22Weak"]),
                                      oracle._p(step["strain12Weak"]), oracle._p(want["divergence"]), oracle._p(want["shear"]),
                                      oracle._p(want["ridgeConvergence"]), oracle._p(want["ridgeShear"]))
    one = np.ones(nC + 1, dtype=np.int32)
    L.orc_principal_stresses_variational(nC, 1, oracle._p(one), oracle._p(step["stress11Weak"]), oracle._p(step["stress22Weak"]),
                                         oracle._p(step["stress12Weak"]), oracle._p(step["replacementPressureWeak"]),
                                         oracle._p(want["principalStress1Weak"]), oracle._p(want["principalStress2Weak"]))
    return want


@pytest.mark.parametrize("path", WEAK_FILES, ids=WEAK_IDS)
def test_oracle_reproduces_the_reference_executed_weak_step(path):
    """config_strain_scheme = config_stress_divergence_scheme = 'weak' through a whole step: the pre-subcycle (its
    weak init_subcycle_variables branch), the weak subcycles, seaice_final_divergence_shear_weak, the weak principal
    stresses and ocean_stress_final, against the arrays the reference's statements wrote."""
    from mpas_seaice_b200 import weakmesh
    mesh, var, opts, cat, forcing, pre, out, nsub, config_dt, prov = _load(path)
    for name in ("seaice_strain_tensor_weak", "seaice_stress_tensor_weak", "seaice_stress_divergence_weak",
                 "seaice_final_divergence_shear_weak", "init_subcycle_variables", "principal_stresses"):
        assert name in prov, name
    nC, nV = mesh.nCells, mesh.nVertices
    step, vm = _oracle_pre(mesh, opts, cat, forcing, pre, config_dt, _first_step_prev(mesh))
    oracle.subcycle_velocity_solver(mesh, dict(var, weak=weakmesh.weak_fields(mesh)), step, opts, nsub)
    want = _weak_post_oracle(mesh, step)
    osu, osv, ocu, ocv, coef = oracle.ocean_stress_final(mesh, step, opts, variational_init.interior_vertex(mesh))
    for k in WEAK_CELL:
        assert np.array_equal(step[k][:nC], out[k][:nC]), k
    for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "principalStress1Weak", "principalStress2Weak"):
        assert np.array_equal(want[k][:nC], out[k][:nC]), k
    for k, got in (("uVelocity", step["uVelocity"]), ("vVelocity", step["vVelocity"]), ("oceanStressU", osu), ("oceanStressV", osv),
                   ("oceanStressCoeff", coef), ("stressDivergenceU", step["stressDivergenceU"]),
                   ("stressDivergenceV", step["stressDivergenceV"])):
        assert np.array_equal(got[:nV][vm], out[k][:nV][vm]), k
    for k, got in (("oceanStressCellU", ocu), ("oceanStressCellV", ocv)):
        assert np.array_equal(got[:nC], out[k][:nC]), k
    assert np.abs(out["stress12Weak"][:nC]).max() > 0 and np.abs(out["ridgeShear"][:nC]).max() > 0
    assert not out["stress11"].any()                    # the variational pool is not touched by a weak run


def _device_step(solver, host, mesh, cat, forcing, pre, out, nsub, start, switches):
    nC, nV = mesh.nCells, mesh.nVertices
    # aggregate_mass_and_area and the Hibler strength on the device, from the category tracers
    solver.aggregate(cat["iceAreaCategory"][:, :, 0].copy(), cat["iceVolumeCategory"][:, :, 0].copy(),
                     cat["snowVolumeCategory"][:, :, 0].copy(), hibler_strength=True)
    agg = solver.fetch_aggregate(ice_pressure=True)
    for k in ("iceAreaCell", "iceVolumeCell", "snowVolumeCell", "totalMassCell"):
        assert np.array_equal(agg[k][:nC], pre[k][:nC]), k
    # The Hibler strength holds the path's one transcendental: P* h exp(-C (1 - a)) (velocity_solver.F:1419-1436).  CUDA's
    # exp() is a 1-ulp function, the reference's is the host libm's: the device value is held to 1 ulp here, and the step
    # continues from the libm value (what a host that keeps ice_strength passes) so that everything after it stays
    # a bit-for-bit comparison.
    state = dict(iceAreaCell=agg["iceAreaCell"], iceVolumeCell=agg["iceVolumeCell"])
    p_host = oracle.hibler_strength_unmasked(state, nC)
    assert np.all(np.abs(agg["icePressure"][:nC] - p_host[:nC]) <= np.spacing(np.abs(p_host[:nC])))
    cells = dict(forcing, iceAreaCellInitial=agg["iceAreaCell"], iceAreaCell=agg["iceAreaCell"],
                 totalMassCell=agg["totalMassCell"], icePressure=p_host)
    solver.pre_subcycle(cells, cold_start=start, **switches)
    got_pre = solver.fetch_pre()
    vm = pre["solveVelocity"][:nV] == 1
    assert np.array_equal(got_pre["solveStress"][:nC], pre["solveStress"][:nC])
    assert np.array_equal(got_pre["solveVelocity"][:nV], pre["solveVelocity"][:nV])
    assert np.array_equal(got_pre["icePressure"][:nC], pre["icePressure"][:nC])
    for k in ("iceAreaVertex", "totalMassVertex", "totalMassVertexfVertex", "airStressVertexU", "airStressVertexV",
              "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "oceanStressV", "uOceanVelocityVertex",
              "vOceanVelocityVertex", "uVelocityInitial", "vVelocityInitial"):
        assert np.array_equal(got_pre[k][:nV][vm], pre[k][:nV][vm]), k
    solver.run_subcycles(nsub)
    got = solver.post_subcycle(names=host.POST_FIELDS_VARIATIONAL)
    inner = solver.fetch(names=("stress11", "stress22", "stress12", "strain11", "strain22", "strain12",
                                "replacementPressure", "stressDivergenceU", "stressDivergenceV"))
    cm = (pre["solveStress"][:nC] == 1)[:, None] & _valid(mesh)
    for k in POST_CELL:
        assert np.array_equal(got[k][:nC], out[k][:nC]), k
    for k, kf in (("principalStress1Var", "principalStress1"), ("principalStress2Var", "principalStress2")):
        assert np.array_equal(got[k][:nC][cm], out[kf][:nC][cm]), k
    for k in ("uVelocity", "vVelocity", "oceanStressU", "oceanStressV", "oceanStressCoeff"):
        assert np.array_equal(got[k][:nV][vm], out[k][:nV][vm]), k
    for k in POST_CELL2D:
        assert np.array_equal(inner[k][:nC][cm], out[k][:nC][cm]), k
    for k in ("stressDivergenceU", "stressDivergenceV"):
        assert np.array_equal(inner[k][:nV][vm], out[k][:nV][vm]), k
