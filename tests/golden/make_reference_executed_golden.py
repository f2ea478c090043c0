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
         "src/shared/mpas_seaice_velocity_solver_constitutive_relation.F",
         "src/shared/mpas_seaice_velocity_solver_variational.F",
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
}


def interpreter(mesh, var, step, opts, nsub):
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
        fa = F.FArray(v)
        I.globals[k.lower()] = fa          # seaice_mesh_pool's module pointers carry the pool arrays' names
        I.pool[k] = fa
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
    g["strainschemetype"] = g["variational_strain_scheme"]
    g["stressdivergenceschemetype"] = g["variational_stress_divergence_scheme"]
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
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh, constitutive_relation_type=cr)
    opts = dict(opts, **extra)
    if cr == "linear":
        # the linear relation leaves the velocities alone (velocity_solver.F:2529-2541): start from a velocity field
        # that is not at rest, the operator test's (square/operators_strain_stress_divergence/create_ics.py:12-18)
        nV = mesh.nVertices
        x, y = mesh.xVertex[:nV] / mesh.Lx, mesh.yVertex[:nV] / mesh.Ly
        step["uVelocity"][:nV] = np.sin(2.0 * np.pi * 2.56 * x) * np.sin(2.0 * np.pi * 2.56 * y)
        step["vVelocity"][:nV] = np.sin(2.0 * np.pi * 2.56 * x) * np.sin(2.0 * np.pi * 2.56 * y)
    work = common.clone_step(step)
    I, domain = interpreter(mesh, var, work, opts, nsub)
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
    return out, called, time.time() - t0


if __name__ == "__main__":
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        data, called, secs = build(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print("%s: %.1f s, %d KB; interpreted: %s" % (name, secs, os.path.getsize(path) // 1024, ", ".join(called)))
