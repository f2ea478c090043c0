"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCE (no Fortran compiler exists in this image):
tests/golden/fortran_subset.py interprets

    subcycle_velocity_solver            src/shared/mpas_seaice_velocity_solver.F:2404-2464
      single_subcycle_velocity_solver   :2478-2592
        seaice_internal_stress          :2606-2863
          seaice_strain_tensor_variational / seaice_average_strains_on_vertex / seaice_stress_tensor_variational /
          seaice_stress_divergence_variational      src/shared/mpas_seaice_velocity_solver_variational.F:575-1184
          seaice_evp_constitutive_relation[_revised] / seaice_linear_constitutive_relation
                                                    src/shared/mpas_seaice_velocity_solver_constitutive_relation.F:178-373
        ocean_stress_coefficient        :2986-3082
        solve_velocity / solve_velocity_revised     :3096-3342
      seaice_set_special_boundaries_velocity[_masks]  src/shared/mpas_seaice_special_boundaries.F

statement by statement from the files under /root/reference, with the module constants (eccentricity, puny, damping ratios,
turning angle, drag coefficient, sea-water density ...) evaluated from their declarations in the same files.  The MPAS
framework calls inside those routines (pool look-ups, timers, the halo exchange of a single block) are mapped onto this
script's arrays or declared no-ops -- see `Interpreter.noop` below; any other unknown call is an error.

Inputs: a mesh of meshgen.py, the basis arrays of the oracle's precompute and a synthetic per-step state (all three are
INPUTS of the subcycle; they are stored in the fixture).  Outputs: what the reference's statements leave in uVelocity,
vVelocity, stress11/22/12, strain11/22/12, replacementPressure, stressDivergenceU/V, oceanStressCoeff after n subcycles.
The files have the layout of make_golden.py's, so tests/test_golden.py replays them through the oracle (CPU) and through
libevp_b200.so (GPU) and demands the same bits; `provenance` says who computed the outputs.

    python tests/golden/make_reference_executed_golden.py        # needs /root/reference; minutes (an interpreter)
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import common  # noqa: E402
import fortran_subset as F  # noqa: E402
from make_golden import MESH_KEYS, VAR_KEYS  # noqa: E402

REF = os.environ.get("MPAS_SEAICE_REFERENCE", "/root/reference")
FILES = ("src/column/constants/cice/ice_constants_colpkg.F90",
         "src/shared/mpas_seaice_constants.F",
         "src/shared/mpas_seaice_mesh.F",
         "src/shared/mpas_seaice_velocity_solver_constitutive_relation.F",
         "src/shared/mpas_seaice_velocity_solver_variational.F",
         "src/shared/mpas_seaice_velocity_solver_weak.F",
         "src/shared/mpas_seaice_special_boundaries.F",
         "src/shared/mpas_seaice_velocity_solver.F")

CASES = {
    # name: (mesh kind, constitutive relation, subcycles, extra options)
    "refexec_hex12_evp_12": ("hex12", "evp", 12, {}),
    "refexec_ico2_evp_10": ("ico2", "evp", 10, {}),
    "refexec_ico2_revised_8": ("ico2", "evp_revised", 8, {}),
    "refexec_quad10_linear_1": ("quad10", "linear", 1, {}),
    "refexec_quad10_evp_avg_6": ("quad10", "evp", 6, {"average_variational_strain": True}),
    "refexec_ico2_evp_lineardrag_6": ("ico2", "evp", 6, {"ocean_stress_type": "linear"}),
    "refexec_hex12_evp_special_boundaries_8": ("hex12", "evp", 8, {"use_special_boundaries_velocity": True}),
    # the piecewise-linear basis (dense gradient arrays: another instantiation of the device's cell kernel), the 'alternate'
    # denominator, no ocean stress, and config_constitutive_relation_type = 'none'
    "refexec_hex12_pwl_alt_evp_8": ("hex12", "evp", 8, {"_basis": "pwl", "_denominator": "alternate"}),
    "refexec_ico2_no_ocean_stress_6": ("ico2", "evp", 6, {"use_ocean_stress": False}),
    "refexec_quad10_none_3": ("quad10", "none", 3, {}),
    # seaice_set_special_boundaries_velocity_masks (special_boundaries.F:345-401): the masks replaced by given ones
    "refexec_hex12_special_boundary_masks_6": ("hex12", "evp", 6, {"use_special_boundaries_velocity_masks": True}),
    # the same inputs as the oracle-made vectors of make_golden.py, at their full length: a whole 120-subcycle dynamics
    # step of the square test case and of the sphere, interpreted (minutes each)
    "refexec_hex20_evp_120": ("hex20", "evp", 120, {}),
    "refexec_ico3_revised_40": ("ico3", "evp_revised", 40, {}),
    "refexec_ico3_evp_120": ("ico3", "evp", 120, {}),
    # the weak operators (src/shared/mpas_seaice_velocity_solver_weak.F) and the weak-strain / variational-divergence mix
    # (interpolate_strains_weak_to_variational, velocity_solver.F:2877-2972)
    "refexec_hex12_weak_evp_8": ("hex12", "evp", 8, {"strain_scheme": "weak", "stress_divergence_scheme": "weak"}),
    "refexec_ico2_weak_evp_6": ("ico2", "evp", 6, {"strain_scheme": "weak", "stress_divergence_scheme": "weak"}),
    "refexec_quad10_weak_revised_5": ("quad10", "evp_revised", 5, {"strain_scheme": "weak", "stress_divergence_scheme": "weak"}),
    "refexec_ico2_weakvar_evp_5": ("ico2", "evp", 5, {"strain_scheme": "weak", "stress_divergence_scheme": "variational"}),
}
CPU_ONLY = ("refexec_hex12_pwl_alt_evp_8", "refexec_ico2_no_ocean_stress_6", "refexec_quad10_none_3",
            "refexec_hex12_special_boundary_masks_6")
WEAK_STATIC = ("verticesOnEdge", "edgesOnVertex", "normalVectorPolygon", "normalVectorTriangle", "latCellRotated", "latVertexRotated")
WEAK_MESH = ("edgesOnCell", "cellsOnEdge", "dvEdge", "dcEdge", "areaTriangle")
WEAK_STATE = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak",
              "replacementPressureWeak", "strain11Vertex", "strain22Vertex", "strain12Vertex")


def interpreter(mesh, var, step, opts, nsub, weak=None):
    I = F.Interpreter(defined=())          # no macros: the plain CPU build (no MPAS_OPENMP, no offload, no CPRINTEL)
    for f in FILES:
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    assert not [p for p in I.pending if p[0] in ("seaicedensityseawater", "seaiceiceoceandragcoefficient")], I.pending
    # framework calls that do nothing on one block without halo
    I.noop |= {"mpas_timer_start", "mpas_timer_stop", "seaice_load_balance_timers", "mpas_log_write",
               "mpas_dmpar_field_halo_exch", "mpas_dmpar_exch_group_full_halo_exch", "mpas_dmpar_exch_group_reuse_halo_exch"}
    nC, nV = mesh.nCells, mesh.nVertices
    arrays = {}
    for k in ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex", "areaCell"):
        arrays[k] = mesh[k]
    for k in VAR_KEYS:
        arrays[k] = var[k]
    for k, v in step.items():
        if isinstance(v, np.ndarray):
            arrays[k] = v
    for k, v in arrays.items():
        if k in WEAK_STATE:
            continue
        fa = F.FArray(v)
        I.globals[k.lower()] = fa          # seaice_mesh_pool's module pointers carry the pool arrays' names
        I.pool[k] = fa
    if weak is not None:
        # the velocity_weak pool (Registry.xml:3700-3745) and the mesh arrays the weak routines fetch themselves
        for k in WEAK_STATIC:
            I.pool[k] = F.FArray(weak[k])
        for k in WEAK_MESH:
            I.pool[k] = F.FArray(mesh[k])
        for k in WEAK_STATE[:7]:
            I.pool[("velocity_weak", k[:-4])] = F.FArray(step[k])
        for k in WEAK_STATE[7:]:
            I.pool[("velocity_weak_variational", k)] = F.FArray(step[k])
        I.pool.update(nEdges=int(mesh.nEdges), on_a_sphere=bool(mesh.on_a_sphere),
                      sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0))
    # dimensions (seaice_mesh_pool: nCells, nVerticesSolve, vertexDegree) and the pool scalars / configs the routines read
    dims = dict(nCells=nC, nVertices=nV, nVerticesSolve=int(opts.get("nVerticesSolve", nV)), vertexDegree=mesh.vertexDegree,
                maxEdges=mesh.maxEdges)
    for k, v in dims.items():
        I.globals[k.lower()] = int(v)
        I.pool[k] = int(v)
    I.pool["elasticTimeStep"] = float(opts["elasticTimeStep"])
    I.pool["dynamicsTimeStep"] = float(opts["dynamicsTimeStep"])
    I.pool["config_elastic_subcycle_number"] = int(nsub)
    I.pool["config_use_ocean_stress"] = bool(opts.get("use_ocean_stress", True))
    I.pool["config_use_halo_exch"] = False
    I.pool["config_use_special_boundaries_velocity"] = bool(opts.get("use_special_boundaries_velocity", False))
    I.pool["config_use_special_boundaries_velocity_masks"] = bool(opts.get("use_special_boundaries_velocity_masks", False))
    # what seaice_init_velocity_solver / seaice_init_evp set from the namelist (velocity_solver.F:168-214,
    # constitutive_relation.F:75-164): module variables
    g = I.globals
    g["strainschemetype"] = g[opts.get("strain_scheme", "variational") + "_strain_scheme"]
    g["stressdivergenceschemetype"] = g[opts.get("stress_divergence_scheme", "variational") + "_stress_divergence_scheme"]
    g["averagevariationalstrains"] = bool(opts.get("average_variational_strain", False))
    g["oceanstresstype"] = g[{"quadratic": "quadratic_ocean_stress", "linear": "linear_ocean_stress"}[opts.get("ocean_stress_type", "quadratic")]]
    g["constitutiverelationtype"] = g[{"evp": "evp_constitutive_relation", "evp_revised": "revised_evp_constitutive_relation",
                                       "linear": "linear_constitutive_relation", "none": "none_constitutive_relation"}[
        opts.get("constitutive_relation_type", "evp")]]
    # seaice_init_special_boundaries points these at the namelist options (special_boundaries.F:27-30, :76-82)
    g["usespecialboundariesvelocity"] = bool(opts.get("use_special_boundaries_velocity", False))
    g["usespecialboundariesvelocitymasks"] = bool(opts.get("use_special_boundaries_velocity_masks", False))
    g["dampingtimescale"] = float(opts["dampingTimescale"])
    g["numericalinertiacoefficient"] = float(opts.get("numericalInertiaCoefficient", 0.0))
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    domain = types.SimpleNamespace(blocklist=block, configs="configs")
    return I, domain


def build(name):
    kind, cr, nsub, extra = CASES[name]
    mesh, var = common.mesh_case(kind, basis=extra.get("_basis", "wachspress"), denominator=extra.get("_denominator", "original"))
    step, opts = common.step_case(mesh, constitutive_relation_type=cr, **({"use_ocean_stress": False} if extra.get("use_ocean_stress") is False else {}))
    opts = dict(opts, **{k: v for k, v in extra.items() if not k.startswith("_")})
    if cr == "linear":
        # the linear relation leaves the velocities alone (velocity_solver.F:2529-2541): start from a velocity field
        # that is not at rest, the operator test's (square/operators_strain_stress_divergence/create_ics.py:12-18)
        nV = mesh.nVertices
        x, y = mesh.xVertex[:nV] / mesh.Lx, mesh.yVertex[:nV] / mesh.Ly
        step["uVelocity"][:nV] = np.sin(2.0 * np.pi * 2.56 * x) * np.sin(2.0 * np.pi * 2.56 * y)
        step["vVelocity"][:nV] = np.sin(2.0 * np.pi * 2.56 * x) * np.sin(2.0 * np.pi * 2.56 * y)
    if opts.get("use_special_boundaries_velocity"):
        # periodic (1), reversed (2) and zero-velocity (3) vertices with chained sources, as tests/test_gpu_parity.py's
        # 1D_velocity_hex-like case: seaice_set_special_boundaries_velocity (special_boundaries.F:253-331) runs before the
        # loop and after every subcycle (velocity_solver.F:2440-2455)
        nV = mesh.nVertices
        vbt, src = np.zeros(nV + 1, dtype=np.int32), np.zeros(nV + 1, dtype=np.int32)
        active = np.nonzero(step["solveVelocity"][:nV] == 1)[0]
        inactive = np.nonzero(step["solveVelocity"][:nV] != 1)[0]
        rng = np.random.default_rng(3)
        pick = rng.choice(inactive, size=min(30, inactive.size), replace=False)
        for i, v in enumerate(pick):
            vbt[v] = 1 + (i % 3)
            if vbt[v] in (1, 2):
                src[v] = int(rng.choice(active)) + 1
        chain = pick[vbt[pick] == 1]
        src[chain[0]] = chain[3] + 1      # a source updated LATER in the sequential loop (the old value is seen)
        src[chain[2]] = chain[1] + 1      # a source updated EARLIER (the new value is seen)
        step["vertexBoundaryType"], step["vertexBoundarySourceLocal"] = vbt, src
    if opts.get("use_special_boundaries_velocity_masks"):
        nC, nV = mesh.nCells, mesh.nVertices
        ss, sv = step["solveStress"].copy(), step["solveVelocity"].copy()
        sv[:nV:7] = 0
        ss[:nC:5] = 0
        step["solveStressSpecialBoundaries"], step["solveVelocitySpecialBoundaries"] = ss, sv
    weak = None
    if opts.get("strain_scheme") == "weak":
        from mpas_seaice_b200 import weakmesh
        weak = weakmesh.weak_fields(mesh)          # INPUTS: the velocity_weak pool's static arrays
        nC, nV = mesh.nCells, mesh.nVertices
        for k in WEAK_STATE:
            step[k] = np.zeros((nV if k.endswith("Vertex") else nC) + 1)
    work = common.clone_step(step)
    I, domain = interpreter(mesh, var, work, opts, nsub, weak)
    t0 = time.time()
    I.call("subcycle_velocity_solver", domain, None)
    called = sorted(set(I.trace))
    out = {"nsub": np.int64(nsub), "provenance": np.array(
        "outputs computed by interpreting the reference's Fortran source (tests/golden/fortran_subset.py): " + ", ".join(called))}
    for k in ("nCells", "nVertices", "maxEdges", "vertexDegree"):
        out["mesh_" + k] = np.int64(mesh[k])
    for k in MESH_KEYS:
        out["mesh_" + k] = mesh[k]
    for k in VAR_KEYS:
        out["var_" + k] = var[k]
    for k, v in step.items():
        if isinstance(v, np.ndarray):
            out["in_" + k] = v
    for k, v in opts.items():
        out["opt_" + k] = np.array(v)
    for k in common.COMPARE_CELL + common.COMPARE_VERTEX:
        out["out_" + k] = work[k]
    if weak is not None:
        for k in WEAK_STATIC:
            out["weak_" + k] = weak[k]
        for k in WEAK_MESH + ("nEdges", "on_a_sphere", "sphere_radius"):
            out["mesh_" + k] = np.array(mesh[k]) if not isinstance(mesh[k], np.ndarray) else mesh[k]
        for k in WEAK_STATE:
            out["out_" + k] = work[k]
    return out, called, time.time() - t0


# ---------------------------------------------------------------------------------------------------------------------
# The init-time precompute: init_velocity_solver_variational_primary_mesh (variational.F:108-344) with everything below it
# -- seaice_calc_variational_metric_terms, seaice_cell_vertices_at_vertex (mesh.F:632), seaice_calc_local_coords
# (variational_shared.F:42-290), seaice_init_velocity_solver_wachspress (wachspress.F:46-1287 incl. the Dunavant tables) or
# seaice_init_velocity_solver_pwl (pwl.F:44-373 with the LU solve of numerics.F), variational_denominator (:358-445).
# Input: the mesh-file arrays; output: the eight static arrays of the velocity_variational pool.
# ---------------------------------------------------------------------------------------------------------------------
INIT_FILES = ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F",
              "src/shared/mpas_seaice_numerics.F", "src/shared/mpas_seaice_mesh.F",
              "src/shared/mpas_seaice_velocity_solver_variational_shared.F", "src/shared/mpas_seaice_velocity_solver_wachspress.F",
              "src/shared/mpas_seaice_velocity_solver_pwl.F", "src/shared/mpas_seaice_velocity_solver_variational.F")
INIT_OUT = ("cellVerticesAtVertex", "tanLatVertexRotatedOverRadius", "basisGradientU", "basisGradientV", "basisIntegralsU",
            "basisIntegralsV", "basisIntegralsMetric", "variationalDenominator")
INIT_MESH = ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex", "edgesOnCell", "xVertex", "yVertex", "zVertex", "xCell", "yCell",
             "zCell", "areaCell", "areaTriangle", "dvEdge")
INIT_CASES = {
    # name: (mesh, basis, denominator, integration type, order)
    "refexec_init_hex5_wachspress": (("planar_hex", 5, 6, 16000.0), "wachspress", "original", "dunavant", 8),
    "refexec_init_ico1_wachspress": (("icosphere", 1), "wachspress", "original", "dunavant", 8),
    "refexec_init_quad4_wachspress_alt": (("planar_quad", 4, 4, 16000.0), "wachspress", "alternate", "dunavant", 4),
    "refexec_init_hex4_trapezoidal": (("planar_hex", 4, 5, 16000.0), "wachspress", "original", "trapezoidal", 3),
    "refexec_init_hex5_pwl": (("planar_hex", 5, 6, 16000.0), "pwl", "original", "dunavant", 8),
    "refexec_init_ico1_pwl": (("icosphere", 1), "pwl", "alternate", "dunavant", 8),
    # the other rules of config_wachspress_integration_type / _order (Registry.xml:603-610)
    "refexec_init_hex3_fekete9": (("planar_hex", 3, 4, 16000.0), "wachspress", "original", "fekete", 9),
    "refexec_init_quad3_dunavant12": (("planar_quad", 3, 3, 16000.0), "wachspress", "alternate", "dunavant", 12),
    "refexec_init_ico1_fekete5": (("icosphere", 1), "wachspress", "original", "fekete", 5),
}
INIT_CPU_ONLY = ("refexec_init_hex3_fekete9", "refexec_init_quad3_dunavant12", "refexec_init_ico1_fekete5")


def init_mesh(spec):
    from mpas_seaice_b200 import meshgen
    return getattr(meshgen, spec[0])(*spec[1:])


def build_init(name):
    from mpas_seaice_b200 import variational_init
    spec, basis, denominator, itype, order = INIT_CASES[name]
    mesh = init_mesh(spec)
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    for f in INIT_FILES:
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    I.noop |= {"mpas_timer_start", "mpas_timer_stop", "mpas_log_write"}
    out = dict(cellVerticesAtVertex=np.zeros((nV + 1, D), np.int32), tanLatVertexRotatedOverRadius=np.zeros(nV + 1),
               variationalDenominator=np.zeros(nV + 1))
    for k in INIT_OUT[2:7]:
        out[k] = np.zeros((nC + 1, M, M))
    arrays = dict(out)
    for k in INIT_MESH:
        arrays[k] = mesh[k]
    arrays["interiorVertex"] = variational_init.interior_vertex(mesh).astype(np.int32)      # boundary pool (mesh.F:423)
    for k, v in arrays.items():
        I.pool[k] = F.FArray(v)
    sphere = bool(mesh.on_a_sphere)
    I.pool.update(on_a_sphere=sphere, sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0), nCells=nC,
                  nVertices=nV, vertexDegree=D, maxEdges=M)
    t0 = time.time()
    # rotateCartesianGrid / includeMetricTerms: the Registry defaults on a sphere (Registry.xml:571-578), off on a plane
    I.call("init_velocity_solver_variational_primary_mesh", "mesh", "velocity_variational", "boundary", sphere, sphere,
           basis, denominator, itype, order)
    called = sorted(set(I.trace))
    data = {"provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): " + ", ".join(called)),
            "spec": np.array(repr(spec)), "basis": np.array(basis), "denominator": np.array(denominator),
            "integration_type": np.array(itype), "integration_order": np.int64(order)}
    for k in INIT_MESH:
        data["mesh_" + k] = mesh[k]
    for k in INIT_OUT:
        data["out_" + k] = out[k]
    return data, called, time.time() - t0


# ---------------------------------------------------------------------------------------------------------------------
# A whole dynamics step, the body of seaice_run_velocity_solver (velocity_solver.F:562-595):
#   velocity_solver_pre_subcycle (:613-671: aggregate_mass_and_area, calculation_masks, new_ice_velocities, ice_strength,
#   air_stress, coriolis_force_coefficient, ocean_stress, surface_tilt, init_subcycle_variables, with
#   seaice_interpolate_cell_to_vertex of mesh.F) -> subcycle_velocity_solver -> velocity_solver_post_subcycle (:3360-3380:
#   final_divergence_shear, principal_stresses_driver, ocean_stress_final with seaice_interpolate_vertex_to_cell).
# Input: the category tracers, the coupler fields and the state carried from the previous step; output: every field the
# three phases leave in the pools.
# ---------------------------------------------------------------------------------------------------------------------
STEP_FILES = ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F",
              "src/shared/mpas_seaice_mesh.F", "src/shared/mpas_seaice_velocity_solver_constitutive_relation.F",
              "src/shared/mpas_seaice_velocity_solver_variational.F", "src/shared/mpas_seaice_special_boundaries.F",
              "src/shared/mpas_seaice_velocity_solver.F")
STEP_CASES = {
    # name: (mesh kind, state kind, config_dt, subcycles, categories)
    "refexec_step_ico2_stateB_6": ("ico2", "B", 3600.0, 6, 1),
    "refexec_step_hex12_square_5": ("hex12", "square", 3600.0, 5, 1),
    "refexec_step_ico2_stateB_3cat_4": ("ico2", "B", 3600.0, 4, 3),
    # two consecutive steps with a MOVING ice edge (caps poleward of 70 N, then of 50 N): the second step starts from the
    # first one's velocities, stresses and solveVelocityPrevious -- new_ice_velocities' branches (:1250-1279)
    "refexec_step_ico2_moving_edge_2x4": ("ico2", "caps:70,50", 3600.0, 4, 1),
    # the namelist switches of the pre-subcycle: config_geostrophic_surface_tilt = false (surface_tilt_ssh_gradient
    # :2024-2170 from the coupler's sea-surface tilt), config_use_air_stress / config_use_surface_tilt = false
    "refexec_step_ico2_sshtilt_3": ("ico2", "B", 3600.0, 3, 1, dict(geostrophic_surface_tilt=False)),
    "refexec_step_ico2_noair_notilt_3": ("ico2", "B", 3600.0, 3, 1, dict(use_air_stress=False, use_surface_tilt=False)),
    # ice shelves: landIceMask = 1 on a patch inside the ice cover; init_ice_shelve_vertex_mask (:481-544) and
    # dynamically_locked_cell_mask (:402-467) run first, the calculation masks (:1023, :1131) then leave the patch out
    "refexec_step_ico2_landice_3": ("ico2", "B", 3600.0, 3, 1, dict(land_ice=True)),
    "refexec_step_hex12_landice_3": ("hex12", "square", 3600.0, 3, 1, dict(land_ice=True)),
    # the other namelist options of the subcycle, carried through a WHOLE step (pre-subcycle -> subcycles -> post-subcycle):
    # revised EVP (init_subcycle_variables' uVelocityInitial feeds solve_velocity_revised), linear ocean drag with
    # averaged variational strains (seaice_average_strains_on_vertex between strain and stress, the final
    # divergence / shear from the unaveraged strains)
    "refexec_step_ico2_revised_4": ("ico2", "B", 3600.0, 4, 1, dict(opts=dict(constitutive_relation_type="evp_revised"))),
    "refexec_step_hex12_lineardrag_avg_4": ("hex12", "square", 3600.0, 4, 1,
                                            dict(opts=dict(ocean_stress_type="linear", average_variational_strain=True))),
    # config_use_ocean_stress = false through the whole step (ocean_stress :1849, ocean_stress_coefficient :3021,
    # ocean_stress_final :3700 all take their 'off' branch), three categories
    # the weak operators through the whole step: init_subcycle_variables' weak branch (:2350-2365), the weak subcycle,
    # seaice_final_divergence_shear_weak (weak.F:651-751) and the weak principal stresses (:3500-3515)
    "refexec_step_ico2_weak_4": ("ico2", "B", 3600.0, 4, 1, dict(opts=dict(strain_scheme="weak", stress_divergence_scheme="weak"))),
    "refexec_step_hex12_weak_3": ("hex12", "square", 3600.0, 3, 1, dict(opts=dict(strain_scheme="weak", stress_divergence_scheme="weak"))),
    # weak strains interpolated to the variational stress points (interpolate_strains_weak_to_variational :2877-2972),
    # variational stress divergence, the variational post-subcycle
    "refexec_step_ico2_weakvar_3": ("ico2", "B", 3600.0, 3, 1, dict(opts=dict(strain_scheme="weak", stress_divergence_scheme="variational"))),
    "refexec_step_ico2_no_ocean_stress_3cat_3": ("ico2", "B", 3600.0, 3, 3, dict(opts=dict(use_ocean_stress=False))),
}
STEP_CPU_ONLY = ("refexec_step_ico2_landice_3", "refexec_step_hex12_landice_3", "refexec_step_ico2_revised_4",
                 "refexec_step_hex12_lineardrag_avg_4", "refexec_step_ico2_no_ocean_stress_3cat_3",
                 "refexec_step_ico2_weak_4", "refexec_step_hex12_weak_3", "refexec_step_ico2_weakvar_3")
STEP_OUT = {
    "velocity_solver": ("solveStress", "solveVelocity", "solveVelocityPrevious", "icePressure", "airStressCellU", "airStressCellV",
                        "uVelocity", "vVelocity", "uVelocityInitial", "vVelocityInitial", "stressDivergenceU", "stressDivergenceV",
                        "oceanStressU", "oceanStressV", "oceanStressCoeff", "airStressVertexU", "airStressVertexV",
                        "surfaceTiltForceU", "surfaceTiltForceV", "uOceanVelocityVertex", "vOceanVelocityVertex",
                        "totalMassVertexfVertex", "divergence", "shear", "oceanStressCellU", "oceanStressCellV"),
    "icestate": ("iceAreaVertex", "totalMassVertex", "totalMassCell", "iceAreaCellInitial"),
    "tracers_aggregate": ("iceAreaCell", "iceVolumeCell", "snowVolumeCell"),
    "velocity_variational": ("strain11", "strain22", "strain12", "stress11", "stress22", "stress12", "replacementPressure",
                             "principalStress1", "principalStress2"),
    "ridging": ("ridgeConvergence", "ridgeShear"),
}


def step_state(mesh, state_kind, n_cat):
    """cell state + forcing of mpas_seaice_b200.synthetic, the ice spread over n_cat thickness categories"""
    from mpas_seaice_b200 import meshgen, synthetic
    if state_kind == "square":
        scaled = meshgen.Mesh(mesh)
        scaled.xCell = mesh.xCell * (1.28e6 / mesh.Lx)
        scaled.yCell = mesh.yCell * (1.28e6 / mesh.Ly)
        state = synthetic.square_state(scaled)
    else:
        state = synthetic.sphere_state(mesh, kind=state_kind)
    nC = mesh.nCells
    w = np.array([0.5, 0.3, 0.2][:n_cat]) if n_cat > 1 else np.array([1.0])
    w = w / w.sum()
    cat = {}
    for k_cell, k_cat in (("iceAreaCell", "iceAreaCategory"), ("iceVolumeCell", "iceVolumeCategory"), ("snowVolumeCell", "snowVolumeCategory")):
        a = np.zeros((nC + 1, n_cat, 1))
        for k in range(n_cat):
            a[:, k, 0] = state[k_cell] * w[k] * (1.0 + 0.1 * k if k_cell != "iceAreaCell" else 1.0)
        cat[k_cat] = a
    return state, cat


def build_step(name):
    from mpas_seaice_b200 import variational_init
    kind, state_kind, config_dt, nsub, n_cat = STEP_CASES[name][:5]
    sw = dict(use_air_stress=True, use_surface_tilt=True, geostrophic_surface_tilt=True)
    sw.update(STEP_CASES[name][5] if len(STEP_CASES[name]) > 5 else {})
    land_ice = bool(sw.pop("land_ice", False))
    opt_over = dict(sw.pop("opts", {}))
    mesh, var = common.mesh_case(kind)
    later = []
    if state_kind.startswith("caps:"):
        lats = [float(x) for x in state_kind[5:].split(",")]
        state, cat = step_state(mesh, "B", n_cat)
        seq = []
        for lat0 in lats:
            cap = (np.degrees(mesh.latCell[:mesh.nCells]) > lat0) | (np.degrees(mesh.latCell[:mesh.nCells]) < -60.0)
            c = {}
            for k_cat, val in (("iceAreaCategory", 1.0), ("iceVolumeCategory", 1.0), ("snowVolumeCategory", 0.0)):
                a = np.zeros((mesh.nCells + 1, 1, 1))
                a[:mesh.nCells, 0, 0] = np.where(cap, val, 0.0)
                c[k_cat] = a
            seq.append(c)
        cat, later = seq[0], seq[1:]
    else:
        state, cat = step_state(mesh, state_kind, n_cat)
    cr = opt_over.get("constitutive_relation_type", "evp")
    _, opts = common.step_case(mesh, constitutive_relation_type=cr)   # elasticTimeStep, dampingTimescale ... as seaice_init_evp sets them
    opts = dict(opts, **opt_over)
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    is_weak = opts.get("strain_scheme", "variational") == "weak"
    for f in STEP_FILES[:-1] + (("src/shared/mpas_seaice_velocity_solver_weak.F",) if is_weak else ()) + STEP_FILES[-1:]:
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    I.noop |= {"mpas_timer_start", "mpas_timer_stop", "mpas_log_write", "seaice_load_balance_timers", "mpas_dmpar_field_halo_exch",
               "mpas_dmpar_exch_group_full_halo_exch", "mpas_dmpar_exch_group_reuse_halo_exch",
               "seaice_mesh_pool_update"}         # seaice_mesh_pool's pointers ARE the pool arrays here
    zc, zv, zcm = (lambda: np.zeros(nC + 1)), (lambda: np.zeros(nV + 1)), (lambda: np.zeros((nC + 1, M)))
    P = {}
    for k, a in cat.items():
        P[("tracers", k)] = a.copy()
    for k in STEP_OUT["tracers_aggregate"]:
        P[("tracers_aggregate", k)] = zc()
    for k in ("totalMassCell", "iceAreaCellInitial", "openWaterArea"):
        P[("icestate", k)] = zc()
    for k in ("iceAreaVertex", "totalMassVertex"):
        P[("icestate", k)] = zv()
    for k in ("uAirVelocity", "vAirVelocity", "airDensity"):
        P[("atmos_coupling", k)] = np.ascontiguousarray(state[k], dtype=np.float64).copy()
    for k in ("uOceanVelocity", "vOceanVelocity"):
        P[("ocean_coupling", k)] = np.ascontiguousarray(state[k], dtype=np.float64).copy()
    for k in ("seaSurfaceTiltU", "seaSurfaceTiltV"):
        P[("ocean_coupling", k)] = zc()
    if not sw["geostrophic_surface_tilt"]:
        P[("ocean_coupling", "seaSurfaceTiltU")][:nC] = 1e-6 * np.sin(3 * mesh.lonCell[:nC]) * np.cos(mesh.latCell[:nC])
        P[("ocean_coupling", "seaSurfaceTiltV")][:nC] = -2e-6 * np.cos(2 * mesh.lonCell[:nC]) * np.cos(mesh.latCell[:nC])
    P[("ocean_coupling", "landIceMask")] = np.zeros(nC + 1, np.int32)
    if land_ice:
        if mesh.on_a_sphere:
            lat, lon = np.degrees(mesh.latCell[:nC]), mesh.lonCell[:nC]
            shelf = (lat < -65.0) | ((lat > 74.0) & (np.cos(lon) > 0.3))
        else:
            shelf = (mesh.xCell[:nC] > 0.62 * mesh.Lx) & (mesh.yCell[:nC] > 0.55 * mesh.Ly)
        assert 0 < shelf.sum() < nC
        P[("ocean_coupling", "landIceMask")][:nC] = shelf
    P[("ocean_coupling", "landIceMaskVertex")] = np.full(nV + 1, -7, np.int32)       # written by the reference below
    P[("velocity_solver", "dynamicallyLockedCellsMask")] = np.full(nC + 1, -7, np.int32)
    P[("boundary", "interiorVertex")] = variational_init.interior_vertex(mesh).astype(np.int32)
    P[("velocity_solver", "solveStress")] = np.zeros(nC + 1, np.int32)
    for k in ("solveVelocity", "solveVelocityPrevious"):
        P[("velocity_solver", k)] = np.zeros(nV + 1, np.int32)         # no Registry default: the first step of a run
    for k in ("icePressure", "airStressCellU", "airStressCellV", "divergence", "shear", "oceanStressCellU", "oceanStressCellV"):
        P[("velocity_solver", k)] = zc()
    for k in ("uVelocity", "vVelocity", "uVelocityInitial", "vVelocityInitial", "stressDivergenceU", "stressDivergenceV",
              "oceanStressU", "oceanStressV", "oceanStressCoeff", "airStressVertexU", "airStressVertexV", "surfaceTiltForceU",
              "surfaceTiltForceV", "seaSurfaceTiltVertexU", "seaSurfaceTiltVertexV", "uOceanVelocityVertex",
              "vOceanVelocityVertex", "totalMassVertexfVertex"):
        P[("velocity_solver", k)] = zv()
    for k in STEP_OUT["velocity_variational"]:
        P[("velocity_variational", k)] = zcm()
    for k in ("strain11", "strain22", "strain12", "stress11", "stress22", "stress12", "replacementPressure", "principalStress1",
              "principalStress2"):
        P[("velocity_weak", k)] = zc()
    for k in STEP_OUT["ridging"]:
        P[("ridging", k)] = zc()
    for k in list(mesh.keys()):
        if isinstance(mesh[k], np.ndarray):
            P[("mesh", k)] = mesh[k]
    for k in VAR_KEYS:
        P[("velocity_variational", k)] = var[k]
    for (pool, k), v in P.items():
        fa = F.FArray(v)
        I.pool[(pool, k)] = fa
        I.pool.setdefault(k, fa)
        if pool != "velocity_weak":
            I.globals[k.lower()] = fa                  # seaice_mesh_pool's module pointers
    if is_weak:
        from mpas_seaice_b200 import weakmesh
        weak = weakmesh.weak_fields(mesh)          # INPUTS: the velocity_weak pool's static arrays (seaice_normal_vectors)
        for k in WEAK_STATIC:
            I.pool[k] = F.FArray(weak[k])
        for k in WEAK_STATE[7:]:
            I.pool[("velocity_weak_variational", k)] = F.FArray(zv())
        I.pool.update(nEdges=int(mesh.nEdges), on_a_sphere=bool(mesh.on_a_sphere),
                      sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0))
    dims = dict(nCells=nC, nVertices=nV, nVerticesSolve=nV, nCellsSolve=nC, vertexDegree=D, maxEdges=M, nCategories=n_cat)
    I.pool.update(dims)
    for k, v in dims.items():
        I.globals[k.lower()] = v
    I.pool.update(config_use_halo_exch=False, config_aggregate_halo_exch=False, config_reuse_halo_exch=False,
                  config_use_column_package=False, config_use_column_vertical_thermodynamics=False,
                  config_use_air_stress=bool(sw["use_air_stress"]), config_use_ocean_stress=bool(opts.get("use_ocean_stress", True)),
                  config_use_surface_tilt=bool(sw["use_surface_tilt"]),
                  config_geostrophic_surface_tilt=bool(sw["geostrophic_surface_tilt"]), config_calc_velocity_masks=True,
                  config_stress_divergence_scheme=str(opts.get("stress_divergence_scheme", "variational")),
                  config_strain_scheme=str(opts.get("strain_scheme", "variational")),
                  config_elastic_subcycle_number=int(nsub), config_use_special_boundaries_velocity=False,
                  config_use_special_boundaries_velocity_masks=False,
                  elasticTimeStep=float(opts["elasticTimeStep"]), dynamicsTimeStep=float(opts["dynamicsTimeStep"]))
    g = I.globals
    g["strainschemetype"] = g[opts.get("strain_scheme", "variational") + "_strain_scheme"]
    g["stressdivergenceschemetype"] = g[opts.get("stress_divergence_scheme", "variational") + "_stress_divergence_scheme"]
    g["averagevariationalstrains"] = bool(opts.get("average_variational_strain", False))
    g["oceanstresstype"] = g[{"quadratic": "quadratic_ocean_stress", "linear": "linear_ocean_stress"}[opts.get("ocean_stress_type", "quadratic")]]
    g["constitutiverelationtype"] = g[{"evp": "evp_constitutive_relation", "evp_revised": "revised_evp_constitutive_relation"}[cr]]
    g["usespecialboundariesvelocity"] = False
    g["usespecialboundariesvelocitymasks"] = False
    g["dampingtimescale"] = float(opts["dampingTimescale"])
    g["numericalinertiacoefficient"] = float(opts.get("numericalInertiaCoefficient", 0.0))
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    domain = types.SimpleNamespace(blocklist=block, configs="configs")
    data = {"nsub": np.int64(nsub), "config_dt": np.float64(config_dt), "kind": np.array(kind), "state_kind": np.array(state_kind)}
    for k, a in cat.items():
        data["in_" + k] = a.copy()
    for k in ("uAirVelocity", "vAirVelocity", "airDensity", "uOceanVelocity", "vOceanVelocity"):
        data["in_" + k] = np.ascontiguousarray(state[k], dtype=np.float64)
    for k, v in opts.items():
        data["opt_" + k] = np.array(v)
    for k, v in sw.items():
        data["sw_" + k] = np.array(bool(v))
    for k in ("seaSurfaceTiltU", "seaSurfaceTiltV"):
        data["in_" + k] = P[("ocean_coupling", k)].copy()
    t0 = time.time()
    # seaice_init_velocity_solver's two mask routines (:237-238)
    if not land_ice:
        P[("ocean_coupling", "landIceMaskVertex")][:] = 0
    else:
        I.call("init_ice_shelve_vertex_mask", domain)
        I.call("dynamically_locked_cell_mask", domain)
        data["in_landIceMask"] = P[("ocean_coupling", "landIceMask")].copy()
        data["ref_landIceMaskVertex"] = P[("ocean_coupling", "landIceMaskVertex")].copy()
        data["ref_dynamicallyLockedCellsMask"] = P[("velocity_solver", "dynamicallyLockedCellsMask")].copy()
        data["ref_interiorVertex"] = P[("boundary", "interiorVertex")].copy()
    I.call("velocity_solver_pre_subcycle", domain)
    for pool, names in STEP_OUT.items():
        for k in names:
            data["pre_" + k] = P[(pool, k)].copy()
    I.call("subcycle_velocity_solver", domain, None)
    # The pre-subcycle ran with config_use_column_package = .false.: the column package is not part of this path, so the
    # categories are aggregated here (:685) and the strength is Hibler's (:1419) instead of colpkg_ice_strength.  The
    # ridging parameters of the post-subcycle (variational.F:1283-1300) exist only with the package on (the Registry
    # default), and nothing else in the post-subcycle reads the switch: on from here.
    I.pool["config_use_column_package"] = True
    I.call("velocity_solver_post_subcycle", domain)
    for pool, names in STEP_OUT.items():
        for k in names:
            data["out_" + k] = P[(pool, k)].copy()
    if is_weak:
        for k in ("strain11", "strain22", "strain12", "stress11", "stress22", "stress12", "replacementPressure", "principalStress1",
                  "principalStress2"):
            data["out_" + k + "Weak"] = P[("velocity_weak", k)].copy()
    for n_step, c in enumerate(later, 2):              # further steps: new tracers in, the dynamic state carried in the pools
        for k, a in c.items():
            P[("tracers", k)][...] = a
            data["in%d_%s" % (n_step, k)] = a.copy()
        I.pool["config_use_column_package"] = False
        I.call("velocity_solver_pre_subcycle", domain)
        for pool, names in STEP_OUT.items():
            for k in names:
                data["pre%d_%s" % (n_step, k)] = P[(pool, k)].copy()
        I.call("subcycle_velocity_solver", domain, None)
        I.pool["config_use_column_package"] = True
        I.call("velocity_solver_post_subcycle", domain)
        for pool, names in STEP_OUT.items():
            for k in names:
                data["out%d_%s" % (n_step, k)] = P[(pool, k)].copy()
    data["n_steps"] = np.int64(1 + len(later))
    called = sorted(set(I.trace))
    data["provenance"] = np.array("outputs computed by interpreting the reference's Fortran source "
                                  "(tests/golden/fortran_subset.py): " + ", ".join(called))
    return data, called, time.time() - t0


# ---------------------------------------------------------------------------------------------------------------------
# The transport row's options: seaice_normal_vectors (mesh.F:703-2007) and config_advection_type = 'upwind'
# (mpas_seaice_advection_upwind.F: define_tracer_connectivities :145 and seaice_run_advection_upwind :385 with everything
# below them, the module's own connectivity table and logical parameters).
# ---------------------------------------------------------------------------------------------------------------------
OPTION_MESHES = {"hex": ("planar_hex", 8, 9, 1000.0), "quad": ("planar_quad", 8, 8, 1000.0), "ico": ("icosphere", 2),
                 "band": ("latlon_band", 24, 10, 60.0)}
UPWIND_NAMES = ("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory", "surfaceTemperature", "iceEnthalpy", "iceSalinity",
                "snowEnthalpy")


def build_normals(kind, remove_metric_terms):
    from mpas_seaice_b200 import irmesh, variational_init
    mesh = init_mesh(OPTION_MESHES[kind])
    irf = irmesh.ir_fields(mesh)
    iv = variational_init.interior_vertex(mesh).astype(np.int32)
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    I.load(os.path.join(REF, "src/shared/mpas_seaice_mesh.F"))
    I.noop |= {"mpas_log_write"}
    for k in list(mesh.keys()):
        if isinstance(mesh[k], np.ndarray):
            I.pool[k] = F.FArray(mesh[k])
    for k in ("verticesOnEdge", "edgesOnVertex", "xEdge", "yEdge", "zEdge"):
        I.pool[k] = F.FArray(irf[k])
    I.pool.update(nCells=nC, nVertices=nV, nVerticesSolve=nV, vertexDegree=D, maxEdges=M, on_a_sphere=bool(mesh.on_a_sphere),
                  sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0))
    out = dict(normalVectorPolygon=np.zeros((nC + 1, M, 2)), normalVectorTriangle=np.zeros((nV + 1, D, 2)),
               latCellRotated=np.zeros(nC + 1), latVertexRotated=np.zeros(nV + 1))
    I.call("seaice_normal_vectors", "mesh", F.FArray(out["normalVectorPolygon"]), F.FArray(out["normalVectorTriangle"]),
           F.FArray(iv), True, bool(remove_metric_terms), F.FArray(out["latCellRotated"]), F.FArray(out["latVertexRotated"]))
    data = {"spec": np.array(repr(OPTION_MESHES[kind])), "remove_metric_terms": np.array(bool(remove_metric_terms)),
            "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))}
    for k in ("xCell", "xVertex"):
        data["mesh_" + k] = mesh[k]
    for k, v in out.items():
        data["out_" + k] = v
    return data


def build_upwind(kind, nsteps=2):
    from mpas_seaice_b200 import irmesh, variational_init, ir_host
    from oracle import ir as oir, upwind as oup          # the oracle only supplies INPUTS here: geometry for the velocity scale
    from test_transport_options import _upwind_state
    from test_oracle_ir import smooth_divergent_velocity
    mesh = init_mesh(OPTION_MESHES[kind])
    irf = irmesh.ir_fields(mesh)
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    nK = 2
    normals = build_normals(kind, False)                                  # normalVectorEdge: the reference's own (:122)
    nve = normals["out_normalVectorPolygon"]
    var = _upwind_state(mesh, np.random.default_rng(11), n_cat=nK, table="reference")
    u, v = smooth_divergent_velocity(mesh, oir.init_geometry(mesh, irf))
    interior = ir_host.interior_edge(mesh)
    I = F.Interpreter(defined=())
    for f in ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F",
              "src/shared/mpas_seaice_advection_upwind.F"):
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    I.noop |= {"mpas_log_write", "mpas_timer_start", "mpas_timer_stop"}
    layers = dict(iceEnthalpy=3, iceSalinity=3, snowEnthalpy=2)
    byname = {x.name: x.array for x in var}
    L1, L2 = {}, {}
    for n in UPWIND_NAMES:
        nl = layers.get(n, 1)
        a = np.zeros((nC + 1, nK, nl))
        if n in byname:
            a[:, :, 0] = byname[n]
        else:
            a[:nC] = 1.5
        L1[n], L2[n] = a, np.full_like(a, 7.0)
        I.pool[("tracers", n, 1)], I.pool[("tracers", n, 2)] = F.FArray(L1[n]), F.FArray(L2[n])
        tend = F.FArray(np.full((nC + 1, nK, nl), 3.0))
        I.pool[("tracer_tendencies", n + "Tend")] = I.pool[("tracer_tendencies", n + "Tend", 1)] = tend
        flux = F.FArray(np.full((nE + 1, nK, nl), 4.0))
        I.pool[("tracer_edge_fluxes", n + "EdgeFlux")] = I.pool[("tracer_edge_fluxes", n + "EdgeFlux", 1)] = flux
        for lvl in (1, 2):
            I.pool[("tracer_conservation", n + "Cons", lvl)] = F.FArray(np.zeros(nK))
    for k in list(mesh.keys()):
        if isinstance(mesh[k], np.ndarray):
            I.pool[k] = F.FArray(mesh[k])
    I.pool["verticesOnEdge"] = F.FArray(irf["verticesOnEdge"])
    I.pool.update(uVelocity=F.FArray(u), vVelocity=F.FArray(v), normalVectorEdge=F.FArray(nve), interiorEdge=F.FArray(interior),
                  dynamicsTimeStep=3600.0, nCells=nC, nCellsSolve=nC, nEdges=nE, nVertices=nV, maxEdges=M, nCategories=nK, ONE=1,
                  config_conservation_check=False)
    I.globals["tracerconnectivities"] = F.FArray(np.array([types.SimpleNamespace() for _ in range(7)], dtype=object))

    def get_field(interp, fr, args):
        pool, key = interp.ev(args[0][1], fr), interp.ev(args[1][1], fr)
        lvl = interp.ev(args[3][1], fr) if len(args) > 3 else 1
        fr.bind(args[2][1][1], types.SimpleNamespace(array=interp.pool[(pool, key, lvl)]))

    def field_info(interp, fr, args):
        fr.bind(args[2][1][1], types.SimpleNamespace(ndims=3))

    def shift(interp, fr, args):                       # MPAS_pool_shift_time_levels: the new level becomes the current one
        for n in UPWIND_NAMES:
            t = L1[n].copy()
            L1[n][...] = L2[n]
            L2[n][...] = t

    I.hooks.update(mpas_pool_get_field=get_field, mpas_pool_get_field_info=field_info, mpas_pool_shift_time_levels=shift,
                   halo_exchange_advection=lambda *a: None)        # one block: nothing to exchange
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    domain = types.SimpleNamespace(blocklist=block, configs="configs")
    data = {"spec": np.array(repr(OPTION_MESHES[kind])), "nsteps": np.int64(nsteps), "dt": np.float64(3600.0),
            "in_uVelocity": u, "in_vVelocity": v, "in_normalVectorEdge": nve, "in_interiorEdge": interior}
    for x in var:
        data["in_" + x.name] = x.array.copy()
    I.call("define_tracer_connectivities", True)
    table = I.globals["tracerconnectivities"].a
    data["table"] = np.array(repr([(t.childtracername.strip(), t.parenttracername.strip()) for t in table if t.defined == 1]))
    for step in range(nsteps):
        I.call("seaice_run_advection_upwind", domain, None)
        for x in var:
            data["out%d_%s" % (step + 1, x.name)] = L1[x.name][:, :, 0].copy()
        for n in ("iceEnthalpy", "iceSalinity", "snowEnthalpy"):
            data["out%d_%s" % (step + 1, n)] = L1[n].copy()
    data["provenance"] = np.array("outputs computed by interpreting the reference's Fortran source "
                                  "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))
    return data


# ---------------------------------------------------------------------------------------------------------------------
# The incremental-remapping transport: seaice_run_advection_incremental_remap (incremental_remap.F:2338-2730) with
# everything below it -- volume <-> thickness, incremental_remap_block (:2740), make_masks, construct_linear_tracer_fields
# (gradients, limiter, barycentres), find_departure_points / _triangles, shift_vertices_of_departure_triangle, the
# quadrature points, integrate_fluxes_over_triangles, compute_mass_tracer_products, update_mass_and_tracers,
# zap_small_mass, and with the optional checks on: tracer_local_min_max, sum_tracers, check_tracer_conservation,
# check_tracer_monotonicity.  The tracer linked list (tracer_type, incremental_remap_tracers.F:26-110) is built here as
# objects holding the arrays; the geometry of the incremental_remap pool is an INPUT (oracle.ir.init_geometry).
# ---------------------------------------------------------------------------------------------------------------------
IR_MESHES = {"hex": ("planar_hex", 6, 7, 1000.0), "quad": ("planar_quad", 6, 6, 1000.0), "ico": ("icosphere", 1),
             "quad16": ("planar_quad", 16, 16, 1000.0)}
IR_CASES = {
    # name: (mesh, categories, ice layers, snow layers, steps, checks)
    "refexec_ir_hex": ("hex", 2, 2, 1, 2, False),
    "refexec_ir_ico": ("ico", 2, 2, 1, 2, False),
    "refexec_ir_quad": ("quad", 2, 2, 1, 2, False),
    "refexec_ir_hex_checks": ("hex", 2, 2, 1, 2, True),
    "refexec_ir_quad_checks": ("quad", 2, 1, 0, 1, True),
    "refexec_ir_quad16_checks": ("quad16", 3, 4, 2, 1, True),       # the reference's monotonicity test fires here (vertexDegree 4)
    "refexec_ir_ico_rotated": ("ico", 2, 2, 1, 2, False, dict(rotate=True)),     # config_rotate_cartesian_grid (Registry default)
    "refexec_ir_hex_3qp": ("hex", 2, 2, 0, 2, False, dict(nqp=3)),               # nQuadPoints = 3 (:780-782, :6598)
}


def build_ir(name):
    from mpas_seaice_b200 import irmesh
    from oracle import ir as oir                  # INPUTS only: the incremental_remap pool's geometry, the test state
    from test_oracle_ir import smooth_divergent_velocity, _random_state
    kind, nK, n_ice, n_snow, nsteps, checks = IR_CASES[name][:6]
    extra = IR_CASES[name][6] if len(IR_CASES[name]) > 6 else {}
    rotate, nQP = bool(extra.get("rotate", False)), int(extra.get("nqp", 6))
    mesh = init_mesh(IR_MESHES[kind])
    irf = irmesh.ir_fields(mesh)
    geom = oir.init_geometry(mesh, irf, rotate=rotate)
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    tracers = _random_state(mesh, np.random.default_rng(21), n_cat=nK, n_ice=n_ice, n_snow=n_snow)
    u, v = smooth_divergent_velocity(mesh, geom)
    I = F.Interpreter(defined=())
    for f in ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F",
              "src/shared/mpas_seaice_advection_incremental_remap.F"):
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    I.noop |= {"mpas_log_write", "mpas_timer_start", "mpas_timer_stop", "seaice_critical_error_write_block",
               "seaice_set_tracer_array_pointers", "seaice_update_tracer_halo", "seaice_load_balance_timers",
               "mpas_dmpar_field_halo_exch", "mpas_dmpar_exch_group_full_halo_exch", "mpas_dmpar_exch_group_reuse_halo_exch"}
    work = [t.array.copy() for t in tracers]
    objs = []
    mk = lambda shape, dt=np.float64: F.FArray(np.zeros(shape, dt))
    for i, t in enumerate(tracers):
        nl = t.array.shape[2]
        o = types.SimpleNamespace(tracername=t.name, parentname="", ndims=2 if nl == 1 else 3, parent=None, next=None,
                                  nextinitial=None, haschild=False, nparents=0, isactive=True)
        if nl == 1:
            o.array2d, o.array3d = F.FArray(work[i][:, :, 0]), None
            for f in ("masstracerproduct2d", "xbarycenter2d", "ybarycenter2d", "center2d", "xgrad2d", "ygrad2d", "localmin2d", "localmax2d"):
                setattr(o, f, mk((nC + 1, nK)))
            o.arraymask2d, o.edgeflux2d, o.trianglevalue2d = mk((nC + 1, nK), np.int64), mk((nE + 1, nK)), mk((nE + 1, 6, nQP, nK))
            o.globalsuminit2d, o.globalsumfinal2d = mk((nK,)), mk((nK,))
        else:
            o.array3d, o.array2d = F.FArray(work[i]), None
            for f in ("masstracerproduct3d", "xbarycenter3d", "ybarycenter3d", "center3d", "xgrad3d", "ygrad3d", "localmin3d", "localmax3d"):
                setattr(o, f, mk((nC + 1, nK, nl)))
            o.arraymask3d, o.edgeflux3d = mk((nC + 1, nK, nl), np.int64), mk((nE + 1, nK, nl))
            o.trianglevalue3d = mk((nE + 1, 6, nQP, nK, nl))
            o.globalsuminit3d, o.globalsumfinal3d = mk((nK, nl)), mk((nK, nl))
        objs.append(o)
    for i, t in enumerate(tracers):
        if t.parent is not None:
            objs[i].parent, objs[t.parent].haschild = objs[t.parent], True
            objs[i].nparents, objs[i].parentname = objs[t.parent].nparents + 1, tracers[t.parent].name
        if i + 1 < len(objs):
            objs[i].next = objs[i + 1]
    P = {k: mesh[k] for k in list(mesh.keys()) if isinstance(mesh[k], np.ndarray)}
    P.update(verticesOnEdge=irf["verticesOnEdge"], edgesOnVertex=irf["edgesOnVertex"], coeffs_reconstruct=irf["coeffs_reconstruct"],
             xEdge=irf["xEdge"], yEdge=irf["yEdge"], zEdge=irf["zEdge"],
             indexToCellID=np.arange(1, nC + 2, dtype=np.int32), indexToVertexID=np.arange(1, nV + 2, dtype=np.int32),
             indexToEdgeID=np.arange(1, nE + 2, dtype=np.int32))
    for k in ("xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge", "remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap",
              "transGlobalToCell", "minLengthEdgesOnVertex"):
        P[k] = geom[k]
    for n in oir.GEOM_NAMES:
        P[n + "AvgCell"] = geom["geomAvg"][n]
    P.update(departurePoint=np.zeros((nV + 1, 2)), xTriangle=np.zeros((nE + 1, 6, nQP)), yTriangle=np.zeros((nE + 1, 6, nQP)),
             iCellTriangle=np.zeros((nE + 1, 6), np.int32), triangleArea=np.zeros((nE + 1, 6)), maskEdge=np.zeros(nE + 1, np.int32),
             maskCell=np.zeros(nC + 1, np.int32), maskCategoryCell=np.zeros((nC + 1, nK), np.int32),
             workCategoryCell=np.zeros((nC + 1, nK)), uVelocity=u, vVelocity=v)
    for k, a in P.items():
        I.pool[k] = F.FArray(a)
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nVerticesSolve=nV, nEdges=nE, nEdgesSolve=nE, maxEdges=M, vertexDegree=D,
                  nCategories=nK, nTriPerEdgeRemap=6, maxCellsPerEdgeRemap=6, maxEdgesPerEdgeRemap=6, maxVerticesPerEdgeRemap=8,
                  nQuadPoints=nQP, on_a_sphere=bool(mesh.on_a_sphere), sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0),
                  config_rotate_cartesian_grid=rotate, config_monotonicity_check=bool(checks), config_conservation_check=bool(checks),
                  config_recover_tracer_means_check=False, config_use_halo_exch=False, dynamicsTimeStep=3600.0)
    g = I.globals
    g["nquadpoints"] = nQP
    w = np.zeros(nQP)
    if nQP == 3:
        w[:] = 1.0 / 3.0                                          # seaice_init_advection_incremental_remap :779-784
    else:
        w[:3], w[3:] = g["w1triangleqp"], g["w2triangleqp"]
    g["weightquadpoint"] = F.FArray(w)
    g["tracershead"] = objs[0]                                    # module variable of ..._incremental_remap_tracers
    g["mpas_dmpar_noerr"] = 0
    for k in ("ctest", "etest", "vtest"):
        g[k] = 1
    for k in ("ctestonproc", "etestonproc", "vtestonproc"):
        g[k] = False
    for k in ("ctestblockid", "etestblockid", "vtestblockid"):
        g[k] = 0
    aborts = []

    def sum_real(interp, fr, args):                               # mpas_dmpar_sum_real on one rank: the local sum
        interp._assign(args[2][1], interp.ev(args[1][1], fr), fr)

    def critical(interp, fr, args):                               # seaice_check_critical_error(domain, abortFlag)
        try:
            aborts.append(bool(interp.ev(args[1][1], fr)))
        except F.FortranError:                                    # the driver's abortFlag (:2398) is only ever SET to .true.
            aborts.append(False)                                  # (:8190, :8540): never assigned = no check has fired

    I.hooks.update(mpas_dmpar_sum_real=sum_real, seaice_check_critical_error=critical)
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None, localblockid=0)
    domain = types.SimpleNamespace(blocklist=block, configs="configs", dminfo=types.SimpleNamespace(my_proc_id=0))
    data = {"spec": np.array(repr(IR_MESHES[kind])), "nsteps": np.int64(nsteps), "dt": np.float64(3600.0), "checks": np.array(bool(checks)),
            "rotate": np.array(rotate), "nqp": np.int64(nQP),
            "in_uVelocity": u, "in_vVelocity": v, "n_tracers": np.int64(len(tracers))}
    for i, t in enumerate(tracers):
        data["in_%d" % i] = t.array.copy()
        data["meta_%d" % i] = np.array(repr((t.name, t.parent, bool(t.volume_like))))
    for step in range(1, nsteps + 1):
        del aborts[:]
        I.call("seaice_run_advection_incremental_remap", domain, None)
        for i in range(len(tracers)):
            data["out%d_%d" % (step, i)] = work[i].copy()
        if checks:
            data["out%d_aborts" % step] = np.array(aborts)       # [after update_mass_and_tracers, conservation, monotonicity]
            for i, o in enumerate(objs):
                sfx = "2d" if o.ndims == 2 else "3d"
                data["out%d_sumInit_%d" % (step, i)] = getattr(o, "globalsuminit" + sfx).a.copy()
                data["out%d_sumFinal_%d" % (step, i)] = getattr(o, "globalsumfinal" + sfx).a.copy()
    data["provenance"] = np.array("outputs computed by interpreting the reference's Fortran source "
                                  "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))
    return data


IR_INIT_MESHES = {"hex": ("planar_hex", 6, 7, 1000.0), "quad": ("planar_quad", 6, 6, 1000.0), "ico": ("icosphere", 1),
                  "band": ("latlon_band", 24, 10, 60.0)}
IR_INIT_CASES = {"refexec_irinit_hex": ("hex", False), "refexec_irinit_quad": ("quad", False), "refexec_irinit_ico": ("ico", False),
                 "refexec_irinit_ico_rotated": ("ico", True), "refexec_irinit_band_rotated": ("band", True)}
IR_INIT_OUT = ("transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap",
               "xVertexOnEdge", "yVertexOnEdge", "minLengthEdgesOnVertex")


def build_ir_init(name):
    """seaice_init_advection_incremental_remap (incremental_remap.F:165-816) as written: rotate_global_vectors,
    define_local_to_global_transformations, get_vertex_on_cell_coordinates, get_geometry_incremental_remap,
    compute_geometric_cell_averages.  Framework calls mapped to no-ops: the halo-layer check, the tracer linked list (built
    by the host), mpas_init_reconstruct (coeffs_reconstruct is the framework's, an input of the transport)."""
    from mpas_seaice_b200 import irmesh
    from oracle import ir as oir
    kind, rotate = IR_INIT_CASES[name]
    mesh = init_mesh(IR_INIT_MESHES[kind])
    irf = irmesh.ir_fields(mesh)
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    for f in ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F",
              "src/shared/mpas_seaice_advection_incremental_remap.F"):
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    I.noop |= {"mpas_log_write", "mpas_timer_start", "mpas_timer_stop", "check_halo_layer_number", "seaice_add_tracers_to_linked_list",
               "seaice_init_update_tracer_halo_exch_group", "mpas_rbf_interp_initialize", "mpas_init_reconstruct",
               "mpas_dmpar_field_halo_exch"}
    P = {k: mesh[k] for k in list(mesh.keys()) if isinstance(mesh[k], np.ndarray)}
    P.update(verticesOnEdge=irf["verticesOnEdge"], edgesOnVertex=irf["edgesOnVertex"], coeffs_reconstruct=irf["coeffs_reconstruct"],
             xEdge=irf["xEdge"], yEdge=irf["yEdge"], zEdge=irf["zEdge"], indexToCellID=np.arange(1, nC + 2, dtype=np.int32),
             indexToVertexID=np.arange(1, nV + 2, dtype=np.int32), indexToEdgeID=np.arange(1, nE + 2, dtype=np.int32))
    out = dict(transGlobalToCell=np.zeros((nC + 1, 3, 3)), xVertexOnCell=np.zeros((nC + 1, M)), yVertexOnCell=np.zeros((nC + 1, M)),
               remapEdge=np.zeros(nE + 1, np.int32), cellsOnEdgeRemap=np.zeros((nE + 1, 6), np.int32),
               edgesOnEdgeRemap=np.zeros((nE + 1, 6), np.int32), xVertexOnEdge=np.zeros((nE + 1, 8)),
               yVertexOnEdge=np.zeros((nE + 1, 8)), minLengthEdgesOnVertex=np.zeros(nV + 1))
    for n in oir.GEOM_NAMES:
        out[n + "AvgCell"] = np.zeros(nC + 1)
    P.update(out)
    P.update(transCellToGlobal=np.zeros((nC + 1, 3, 3)), transVertexToGlobal=np.zeros((nV + 1, 3, 3)),
             transGlobalToVertex=np.zeros((nV + 1, 3, 3)), transEdgeToGlobal=np.zeros((nE + 1, 3, 3)),
             transGlobalToEdge=np.zeros((nE + 1, 3, 3)))
    for pre, n in (("Cell", nC), ("Vertex", nV), ("Edge", nE)):
        for ax in "xyz":
            P[ax + pre + "Rotate"] = np.zeros(n + 1)
    for k, a in P.items():
        I.pool[k] = F.FArray(a)
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nVerticesSolve=nV, nEdges=nE, nEdgesSolve=nE, maxEdges=M, vertexDegree=D,
                  nQuadPoints=6, nTriPerEdgeRemap=6, maxCellsPerEdgeRemap=6, maxEdgesPerEdgeRemap=6, maxVerticesPerEdgeRemap=8,
                  on_a_sphere=bool(mesh.on_a_sphere), sphere_radius=float(getattr(mesh, "sphere_radius", 0.0) or 0.0),
                  config_rotate_cartesian_grid=bool(rotate))
    I.globals["tracershead"] = None
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None, localblockid=0)
    domain = types.SimpleNamespace(blocklist=block, configs="configs", dminfo=types.SimpleNamespace(my_proc_id=0))
    I.call("seaice_init_advection_incremental_remap", domain)
    data = {"spec": np.array(repr(IR_INIT_MESHES[kind])), "rotate": np.array(bool(rotate)), "mesh_xCell": mesh.xCell,
            "weightQuadPoint": I.globals["weightquadpoint"].a.copy(),
            "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))}
    for k, v in out.items():
        data["out_" + k] = v
    return data


def build_init_evp():
    """seaice_init_evp (constitutive_relation.F:75-164): the relation type from the namelist string, dampingTimescale,
    numericalInertiaCoefficient (Bouillon et al. 2013) from the time steps and the shortest edge of the mesh."""
    rows = []
    for kind in ("hex", "ico", "band"):
        mesh = init_mesh(OPTION_MESHES[kind])
        for cr, dt_dyn, nsub in (("evp", 3600.0, 120), ("evp_revised", 900.0, 60), ("linear", 120.0, 120), ("none", 1800.0, 1)):
            I = F.Interpreter(defined=())
            I.load(os.path.join(REF, "src/shared/mpas_seaice_velocity_solver_constitutive_relation.F"))
            I.resolve_constants()
            I.noop |= {"mpas_log_write"}
            I.hooks["mpas_dmpar_min_real"] = lambda interp, fr, args: interp._assign(args[2][1], interp.ev(args[1][1], fr), fr)
            I.pool.update(config_constitutive_relation_type=cr, dynamicsTimeStep=dt_dyn, elasticTimeStep=dt_dyn / nsub,
                          nEdgesSolve=int(mesh.nEdges), dvEdge=F.FArray(mesh.dvEdge))
            block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
            I.call("seaice_init_evp", types.SimpleNamespace(blocklist=block, configs="configs", dminfo=None))
            g = I.globals
            rows.append((kind, cr, dt_dyn, nsub, int(g["constitutiverelationtype"]), float(g["dampingtimescale"]),
                         float(g["numericalinertiacoefficient"]), float(mesh.dvEdge[:mesh.nEdges].min())))
    return {"rows": np.array([repr(r) for r in rows]),
            "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): seaice_init_evp")}


def build_weak_post():
    """seaice_final_divergence_shear_weak (weak.F:654-751), as written: `Delta = sqrt(...)` assigns the whole work array
    inside the cell loop, so the ridging shear of every cell is computed from the LAST owned cell's Delta."""
    rng = np.random.default_rng(7)
    nC = 57
    strain = [np.zeros(nC + 1) for _ in range(3)]
    for a in strain:
        a[:nC] = rng.normal(scale=1e-6, size=nC)
    out = {k: np.zeros(nC + 1) for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear")}
    I = F.Interpreter(defined=())
    I.load(os.path.join(REF, "src/shared/mpas_seaice_velocity_solver_constitutive_relation.F"))
    I.load(os.path.join(REF, "src/shared/mpas_seaice_velocity_solver_weak.F"))
    I.resolve_constants()
    for k, a in zip(("strain11", "strain22", "strain12"), strain):
        I.pool[("velocity_weak", k)] = F.FArray(a)
    for k, a in out.items():
        I.pool[k] = F.FArray(a)
    I.pool.update(nCellsSolve=nC, nEdgesOnCell=F.FArray(np.full(nC + 1, 6, np.int32)), config_use_column_package=True)
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    I.call("seaice_final_divergence_shear_weak", block)
    data = {"nCells": np.int64(nC), "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                                          "(tests/golden/fortran_subset.py): seaice_final_divergence_shear_weak")}
    for k, a in zip(("strain11", "strain22", "strain12"), strain):
        data["in_" + k] = a
    for k, a in out.items():
        data["out_" + k] = a
    return data


def build_square_testcase():
    """The synthetic workload of BASELINE.json configs[1] ("square testcase"): init_square_test_case_state / _atmos / _ocean
    (src/shared/mpas_seaice_testing.F:304-343, 357-422, 436-525) on the cell centres of a planar hex mesh scaled to the
    test case's 1.28e6 m domain, time = 0."""
    from mpas_seaice_b200 import meshgen
    mesh = meshgen.planar_hex(12, 14, 16000.0)
    nC, nK = mesh.nCells, 1
    x = mesh.xCell * (1.28e6 / mesh.Lx)
    y = mesh.yCell * (1.28e6 / mesh.Ly)
    I = F.Interpreter(defined=())
    for f in ("src/column/constants/cice/ice_constants_colpkg.F90", "src/shared/mpas_seaice_constants.F", "src/shared/mpas_seaice_testing.F"):
        I.load(os.path.join(REF, f))
    I.resolve_constants()
    out = {k: np.zeros(nC + 1) for k in ("uAirVelocity", "vAirVelocity", "airDensity", "uOceanVelocity", "vOceanVelocity")}
    cat = {k: np.zeros((nC + 1, nK, 1)) for k in ("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory")}
    fa = lambda a: F.FArray(a)
    I.call("init_square_test_case_atmos", fa(out["uAirVelocity"]), fa(out["vAirVelocity"]), fa(out["airDensity"]), fa(x), fa(y), 0.0)
    I.call("init_square_test_case_ocean", fa(out["uOceanVelocity"]), fa(out["vOceanVelocity"]), fa(x), fa(y))
    I.pool.update(nCells=nC, nVertices=int(mesh.nVertices), xCell=fa(x), interiorVertex=fa(np.zeros(mesh.nVertices + 1, np.int32)))
    for k, a in cat.items():
        I.pool[("tracers", k, 1)] = fa(a)
    for k in ("uVelocity", "vVelocity", "uOceanVelocityVertex", "vOceanVelocityVertex"):
        I.pool[k] = fa(np.zeros(mesh.nVertices + 1))
    I.pool["uOceanVelocity"], I.pool["vOceanVelocity"] = fa(out["uOceanVelocity"]), fa(out["vOceanVelocity"])
    I.call("init_square_test_case_state", "mesh", "tracers", "velocity_solver", "ocean_coupling", "boundary")
    data = {"in_x": x, "in_y": y, "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                                        "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))}
    for k, a in out.items():
        data["out_" + k] = a
    for k, a in cat.items():
        data["out_" + k] = a
    return data


def build_boundary(kind):
    """init_boundary (mesh.F:372-630): interiorVertex, interiorCell, interiorEdge -- the integer maps the solver's masks and
    the upwind fluxes are built on."""
    mesh = init_mesh(OPTION_MESHES[kind])
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    I.load(os.path.join(REF, "src/shared/mpas_seaice_mesh.F"))
    I.noop |= {"mpas_dmpar_field_halo_exch", "mpas_log_write"}
    out = dict(interiorVertex=np.zeros(nV + 1, np.int32), interiorCell=np.zeros(nC + 1, np.int32), interiorEdge=np.zeros(nE + 1, np.int32),
               blockIDout=np.zeros(nC + 1, np.int32))
    for k in ("nEdgesOnCell", "cellsOnCell", "cellsOnVertex", "cellsOnEdge"):
        I.pool[k] = F.FArray(mesh[k])
    for k, v in out.items():
        I.pool[k] = F.FArray(v)
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nVerticesSolve=nV, nEdges=nE, nEdgesSolve=nE, vertexDegree=D, maxEdges=M)
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None, blockid=0)
    I.call("init_boundary", types.SimpleNamespace(blocklist=block, configs="configs"))
    data = {"spec": np.array(repr(OPTION_MESHES[kind])), "mesh_xCell": mesh.xCell,
            "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))}
    for k in ("interiorVertex", "interiorCell", "interiorEdge"):
        data["out_" + k] = out[k]
    return data


def culled_quad_mesh():
    """planar_quad 7 x 6 with two columns of cells culled in the lower rows (their references in cellsOnVertex /
    cellsOnCell replaced by nCells+1, what the mesh culler leaves): a channel one cell wide whose cells have no
    interior vertex."""
    from mpas_seaice_b200 import meshgen
    spec = ("planar_quad", 7, 6, 1000.0)
    mesh = meshgen.Mesh(init_mesh(spec))
    nC = mesh.nCells
    ix = np.rint(mesh.xCell[:nC] / 1000.0 - 0.5).astype(int)
    iy = np.rint(mesh.yCell[:nC] / 1000.0 - 0.5).astype(int)
    culled = np.flatnonzero(((ix == 2) | (ix == 4)) & (iy < 4)) + 1
    for k in ("cellsOnVertex", "cellsOnCell"):
        a = mesh[k].copy()
        a[np.isin(a, culled)] = nC + 1
        mesh[k] = a
    return spec, mesh, culled


def build_locked_cells():
    """interior_vertices (mesh.F:423-488) then dynamically_locked_cell_mask (velocity_solver.F:402-467) on the culled mesh"""
    spec, mesh, culled = culled_quad_mesh()
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    I = F.Interpreter(defined=())
    I.load(os.path.join(REF, "src/shared/mpas_seaice_mesh.F"))
    I.load(os.path.join(REF, "src/shared/mpas_seaice_velocity_solver.F"))
    I.noop |= {"mpas_dmpar_field_halo_exch", "mpas_log_write"}
    out = dict(interiorVertex=np.full(nV + 1, -7, np.int32), dynamicallyLockedCellsMask=np.full(nC + 1, -7, np.int32))
    for k in ("nEdgesOnCell", "cellsOnCell", "cellsOnVertex", "verticesOnCell"):
        I.pool[k] = F.FArray(mesh[k])
    for k, v in out.items():
        I.pool[k] = F.FArray(v)
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nVerticesSolve=nV, nEdges=nE, nEdgesSolve=nE, vertexDegree=D, maxEdges=M)
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None, blockid=0)
    domain = types.SimpleNamespace(blocklist=block, configs="configs")
    I.call("interior_vertices", "mesh", "boundary")
    I.call("dynamically_locked_cell_mask", domain)
    return {"spec": np.array(repr(spec)), "culled": culled.astype(np.int32), "out_interiorVertex": out["interiorVertex"],
            "out_dynamicallyLockedCellsMask": out["dynamicallyLockedCellsMask"],
            "provenance": np.array("outputs computed by interpreting the reference's Fortran source "
                                   "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))}


def build_special_boundaries_init():
    """seaice_init_special_boundaries (special_boundaries.F:60-250): vertexBoundarySourceLocal / tracerBoundarySourceLocal
    from the global IDs of the stream arrays, on a block whose local numbering is a permutation of the global one; then
    seaice_set_special_boundaries_tracers (:415-485) on category tracers (zeroed cells, copied cells, a copied cell whose
    source was itself changed earlier in the loop)."""
    mesh = init_mesh(("planar_hex", 6, 7, 1000.0))
    nC, nV = mesh.nCells, mesh.nVertices
    rng = np.random.default_rng(8)
    idV = np.zeros(nV + 1, np.int32)
    idV[:nV] = rng.permutation(nV) + 1
    idC = np.zeros(nC + 1, np.int32)
    idC[:nC] = rng.permutation(nC) + 1
    vType, vSrc = np.zeros(nV + 1, np.int32), np.zeros(nV + 1, np.int32)
    pick = rng.choice(nV, size=18, replace=False)
    vType[pick] = np.tile([1, 2, 3], 6)
    vSrc[pick] = idV[rng.choice(nV, size=18)]                       # global IDs, as the stream delivers them
    cType, cSrc = np.zeros(nC + 1, np.int32), np.zeros(nC + 1, np.int32)
    pickc = np.sort(rng.choice(nC, size=10, replace=False))
    cType[pickc] = np.tile([1, 2], 5)
    cSrc[pickc] = idC[rng.choice(nC, size=10)]
    cSrc[pickc[-1]] = idC[pickc[1]]                                 # a SET cell whose source is an earlier SET cell
    cType[pickc[-1]] = 2
    cSrc[pickc[-2]] = idC[pickc[0]]                                 # ... and one whose source is an earlier ZERO cell
    cType[pickc[-2]] = 2
    vLoc, cLoc = np.full(nV + 1, -7, np.int32), np.full(nC + 1, -7, np.int32)
    nK = 3
    tr = {k: rng.uniform(0.1, 1.0, size=(nC + 1, nK, 1)) for k in ("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory")}
    data = {"in_" + k: v.copy() for k, v in tr.items()}
    I = F.Interpreter(defined=())
    I.load(os.path.join(REF, "src/shared/mpas_seaice_special_boundaries.F"))
    I.noop |= {"mpas_log_write"}
    for k, v in dict(vertexBoundaryType=vType, vertexBoundarySource=vSrc, vertexBoundarySourceLocal=vLoc, indexToVertexID=idV,
                     tracerBoundaryType=cType, tracerBoundarySource=cSrc, tracerBoundarySourceLocal=cLoc, indexToCellID=idC,
                     **tr).items():
        I.pool[k] = F.FArray(v)
    I.pool.update(nCells=nC, nVertices=nV, config_use_special_boundaries_velocity=True,
                  config_use_special_boundaries_velocity_masks=False, config_use_special_boundaries_tracers=True)
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    domain = types.SimpleNamespace(blocklist=block, configs="configs")
    for k in ("usespecialboundariesvelocity", "usespecialboundariesvelocitymasks", "usespecialboundariestracers"):
        I.globals[k] = False                  # the module's logical pointers: the init routines point them at the configs
    I.call("seaice_init_special_boundaries", domain)
    assert I.globals["usespecialboundariesvelocity"] is True and I.globals["usespecialboundariestracers"] is True
    I.call("seaice_set_special_boundaries_tracers", domain)
    data.update(indexToVertexID=idV, vertexBoundaryType=vType, vertexBoundarySource=vSrc, out_vertexBoundarySourceLocal=vLoc,
                indexToCellID=idC, tracerBoundaryType=cType, tracerBoundarySource=cSrc, out_tracerBoundarySourceLocal=cLoc)
    for k, v in tr.items():
        data["out_" + k] = v
    data["provenance"] = np.array("outputs computed by interpreting the reference's Fortran source "
                                  "(tests/golden/fortran_subset.py): " + ", ".join(sorted(set(I.trace))))
    return data


if __name__ == "__main__":
    only = sys.argv[1:]
    os.makedirs(os.path.join(HERE, "options"), exist_ok=True)
    if not only or "refexec_special_boundaries_init" in only:
        np.savez_compressed(os.path.join(HERE, "cpu", "refexec_special_boundaries_init.npz"), **build_special_boundaries_init())
        print("refexec_special_boundaries_init", flush=True)
    if not only or "refexec_locked_cells" in only:
        np.savez_compressed(os.path.join(HERE, "cpu", "refexec_locked_cells.npz"), **build_locked_cells())
        print("refexec_locked_cells", flush=True)
    if not only or "refexec_square_testcase" in only:
        np.savez_compressed(os.path.join(HERE, "options", "refexec_square_testcase.npz"), **build_square_testcase())
        print("refexec_square_testcase", flush=True)
    if not only or "refexec_weak_post" in only:
        np.savez_compressed(os.path.join(HERE, "options", "refexec_weak_post.npz"), **build_weak_post())
        print("refexec_weak_post", flush=True)
    if not only or "refexec_init_evp" in only:
        np.savez_compressed(os.path.join(HERE, "options", "refexec_init_evp.npz"), **build_init_evp())
        print("refexec_init_evp", flush=True)
    for kind in OPTION_MESHES:
        name = "refexec_boundary_%s" % kind
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, "options", name + ".npz"), **build_boundary(kind))
        print(name, flush=True)
    os.makedirs(os.path.join(HERE, "ir"), exist_ok=True)
    for name in IR_INIT_CASES:
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, "ir", name + ".npz"), **build_ir_init(name))
        print(name, flush=True)
    for name in IR_CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        data = build_ir(name)
        np.savez_compressed(os.path.join(HERE, "ir", name + ".npz"), **data)
        print("%s: %.1f s" % (name, time.time() - t0), flush=True)
    os.makedirs(os.path.join(HERE, "options"), exist_ok=True)
    for kind in OPTION_MESHES:
        for rm in (True, False):
            name = "refexec_normals_%s_%s" % (kind, "nometric" if rm else "metric")
            if only and name not in only:
                continue
            np.savez_compressed(os.path.join(HERE, "options", name + ".npz"), **build_normals(kind, rm))
            print(name, flush=True)
    for kind in ("hex", "quad", "ico"):
        name = "refexec_upwind_%s" % kind
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, "options", name + ".npz"), **build_upwind(kind))
        print(name, flush=True)
    for name in STEP_CASES:
        if only and name not in only:
            continue
        data, called, secs = build_step(name)
        path = os.path.join(HERE, "cpu" if name in STEP_CPU_ONLY else "step", name + ".npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **data)
        print("%s: %.1f s, %d KB; interpreted: %s" % (name, secs, os.path.getsize(path) // 1024, ", ".join(called)), flush=True)
    for name in INIT_CASES:
        if only and name not in only:
            continue
        data, called, secs = build_init(name)
        path = os.path.join(HERE, "cpu" if name in INIT_CPU_ONLY else "init", name + ".npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **data)
        print("%s: %.1f s, %d KB; interpreted: %s" % (name, secs, os.path.getsize(path) // 1024, ", ".join(called)), flush=True)
    for name in CASES:
        if only and name not in only:
            continue
        data, called, secs = build(name)
        # fixtures added after this round's GPU time was spent are replayed through the oracle only (tests/golden/cpu/);
        # the device runs the same configurations against the oracle in tests/test_gpu_parity.py
        path = os.path.join(HERE, "cpu" if name in CPU_ONLY else "", name + ".npz")
        np.savez_compressed(path, **data)
        print("%s: %.1f s, %d KB; interpreted: %s" % (name, secs, os.path.getsize(path) // 1024, ", ".join(called)))
