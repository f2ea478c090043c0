"""ctypes front-end of the CPU oracle (oracle/evp_oracle.c, oracle/evp_precompute_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; the product never does.

Parity status: the reference Fortran cannot be built here and stores no outputs, so this oracle
is pinned by the reference's analytic known answers only (tests/test_oracle_kat.py) --
"parity unpinned" against golden outputs of the reference itself.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

EVP, EVP_REVISED, LINEAR, NONE = 1, 2, 3, 4
QUADRATIC_OCEAN_STRESS, LINEAR_OCEAN_STRESS = 1, 2


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("evp_oracle.c", "evp_precompute_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_damping_timescale.restype = C.c_double
        _lib.orc_damping_timescale.argtypes = [C.c_double]
        _lib.orc_numerical_inertia_coefficient.restype = C.c_double
        _lib.orc_numerical_inertia_coefficient.argtypes = [C.c_double, C.c_double]
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _i(x):
    return C.c_int(int(x))


def _d(x):
    return C.c_double(float(x))


# ---------------------------------------------------------------------------------------------
# precompute
# ---------------------------------------------------------------------------------------------

def init_variational(mesh, basis="wachspress", denominator="original", rotate=None, metric=None,
                     integration_type="dunavant", integration_order=8):
    """seaice_init_velocity_solver_variational (variational.F:53-344) on a meshgen.Mesh.
    Returns a dict with the Registry ``velocity_variational`` static fields + interiorVertex."""
    L = lib()
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    on_sphere = bool(mesh.on_a_sphere)
    if rotate is None:
        rotate = on_sphere          # Registry default config_rotate_cartesian_grid = true (Registry.xml:571)
    if metric is None:
        metric = on_sphere          # planar operator tests must switch it off
    out = {}
    tan = np.zeros(nV + 1)
    L.orc_calc_variational_metric_terms(_p(tan), _i(nV), _p(mesh.xVertex), _p(mesh.yVertex), _p(mesh.zVertex),
                                        _d(mesh.sphere_radius), _i(rotate), _i(metric))
    cvav = np.zeros((nV + 1, D), dtype=np.int32)
    L.orc_cell_vertices_at_vertex(_p(cvav), _i(nV), _i(D), _i(M), _p(mesh.nEdgesOnCell), _p(mesh.verticesOnCell),
                                  _p(mesh.cellsOnVertex))
    xl = np.zeros((nC + 1, M))
    yl = np.zeros((nC + 1, M))
    L.orc_calc_local_coords(_p(xl), _p(yl), _i(nC), _i(M), _p(mesh.nEdgesOnCell), _p(mesh.verticesOnCell),
                            _p(mesh.xVertex), _p(mesh.yVertex), _p(mesh.zVertex),
                            _p(mesh.xCell), _p(mesh.yCell), _p(mesh.zCell), _i(rotate), _i(on_sphere))
    GU, GV, SU, SV, SM = (np.zeros((nC + 1, M, M)) for _ in range(5))
    if basis == "wachspress":
        itype = {"dunavant": 0, "trapezoidal": 1}[integration_type]
        err = L.orc_init_velocity_solver_wachspress(_i(nC), _i(M), _p(mesh.nEdgesOnCell), _p(xl), _p(yl),
                                                    _i(itype), _i(integration_order),
                                                    _p(GU), _p(GV), _p(SU), _p(SV), _p(SM))
        assert err == 0, f"wachspress init failed ({err})"
    elif basis == "pwl":
        err = L.orc_init_velocity_solver_pwl(_i(nC), _i(M), _p(mesh.nEdgesOnCell), _p(mesh.edgesOnCell),
                                             _p(mesh.dvEdge), _p(mesh.areaCell), _p(xl), _p(yl),
                                             _p(GU), _p(GV), _p(SM), _p(SU), _p(SV))
        assert err == 0
    elif basis != "none":
        raise ValueError(basis)
    interior = np.zeros(nV + 1, dtype=np.int32)
    L.orc_interior_vertices(_p(interior), _i(nV), _i(D), _i(nC), _p(mesh.cellsOnVertex))
    den = np.zeros(nV + 1)
    L.orc_variational_denominator(_i(nV), _i(D), _i(M), _p(mesh.nEdgesOnCell), _p(mesh.areaTriangle),
                                  _p(mesh.cellsOnVertex), _p(cvav), _p(SM),
                                  _i({"original": 0, "alternate": 1}[denominator]), _p(den))
    out.update(tanLatVertexRotatedOverRadius=tan, cellVerticesAtVertex=cvav, xLocal=xl, yLocal=yl,
               basisGradientU=GU, basisGradientV=GV, basisIntegralsU=SU, basisIntegralsV=SV,
               basisIntegralsMetric=SM, interiorVertex=interior, variationalDenominator=den)
    return out


def integration_factors(integration_type="dunavant", order=8):
    L = lib()
    n = C.c_int(0)
    norm = C.c_double(0)
    u, v, w = np.zeros(512), np.zeros(512), np.zeros(512)
    err = L.orc_get_integration_factors(_i({"dunavant": 0, "trapezoidal": 1}[integration_type]), _i(order),
                                        C.byref(n), _p(u), _p(v), _p(w), C.byref(norm))
    assert err == 0
    return u[:n.value].copy(), v[:n.value].copy(), w[:n.value].copy(), norm.value


# ---------------------------------------------------------------------------------------------
# subcycle
# ---------------------------------------------------------------------------------------------

class _SubcycleArgs(C.Structure):
    _ints = ["nCells", "nVertices", "nVerticesSolve", "maxEdges", "vertexDegree"]
    _fields_ = (
        [(n, C.c_int) for n in _ints]
        + [(n, C.c_void_p) for n in ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex", "cellVerticesAtVertex",
                                     "basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV",
                                     "basisIntegralsMetric", "tanLatVertexRotatedOverRadius",
                                     "variationalDenominator", "areaCell")]
        + [(n, C.c_int) for n in ("constitutiveRelationType", "oceanStressType", "useOceanStress",
                                  "averageVariationalStrains", "useSpecialBoundariesVelocity",
                                  "useSpecialBoundariesVelocityMasks")]
        + [(n, C.c_double) for n in ("elasticTimeStep", "dynamicsTimeStep", "dampingTimescale",
                                     "numericalInertiaCoefficient")]
        + [(n, C.c_void_p) for n in ("solveStress", "solveVelocity", "vertexBoundaryType",
                                     "vertexBoundarySourceLocal", "solveStressSpecialBoundaries",
                                     "solveVelocitySpecialBoundaries",
                                     "icePressure", "totalMassVertex", "totalMassVertexfVertex", "iceAreaVertex",
                                     "airStressVertexU", "airStressVertexV", "surfaceTiltForceU", "surfaceTiltForceV",
                                     "oceanStressU", "oceanStressV", "uOceanVelocityVertex", "vOceanVelocityVertex",
                                     "uVelocityInitial", "vVelocityInitial",
                                     "uVelocity", "vVelocity", "stress11", "stress22", "stress12",
                                     "strain11", "strain22", "strain12", "replacementPressure",
                                     "stressDivergenceU", "stressDivergenceV", "oceanStressCoeff")]
    )


_STATIC = ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex", "cellVerticesAtVertex", "basisGradientU",
           "basisGradientV", "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric",
           "tanLatVertexRotatedOverRadius", "variationalDenominator", "areaCell")
_STEP = ("solveStress", "solveVelocity", "vertexBoundaryType", "vertexBoundarySourceLocal",
         "solveStressSpecialBoundaries", "solveVelocitySpecialBoundaries", "icePressure", "totalMassVertex",
         "totalMassVertexfVertex", "iceAreaVertex", "airStressVertexU", "airStressVertexV", "surfaceTiltForceU",
         "surfaceTiltForceV", "oceanStressU", "oceanStressV", "uOceanVelocityVertex", "vOceanVelocityVertex",
         "uVelocityInitial", "vVelocityInitial", "uVelocity", "vVelocity", "stress11", "stress22", "stress12",
         "strain11", "strain22", "strain12", "replacementPressure", "stressDivergenceU", "stressDivergenceV",
         "oceanStressCoeff")


def subcycle_velocity_solver(mesh, var, step, opts, n_subcycles):
    """subcycle_velocity_solver (velocity_solver.F:2404-2464) on host arrays, IN PLACE on ``step``.

    mesh: meshgen.Mesh; var: dict from init_variational; step: dict of per-step fields (see
    mpas_seaice_b200.synthetic.pre_subcycle); opts: dict with constitutive_relation_type ('evp' |
    'evp_revised' | 'linear' | 'none'), ocean_stress_type, use_ocean_stress, average_variational_strain,
    elasticTimeStep, dynamicsTimeStep, dampingTimescale, numericalInertiaCoefficient.
    """
    L = lib()
    a = _SubcycleArgs()
    a.nCells, a.nVertices = mesh.nCells, mesh.nVertices
    a.nVerticesSolve = int(opts.get("nVerticesSolve", mesh.nVertices))
    a.maxEdges, a.vertexDegree = mesh.maxEdges, mesh.vertexDegree
    keep = []
    for name in _STATIC:
        arr = var[name] if name in var else mesh[name]
        keep.append(arr)
        setattr(a, name, arr.ctypes.data)
    for name in _STEP:
        arr = step.get(name)
        if arr is None:
            setattr(a, name, None)
        else:
            assert arr.flags["C_CONTIGUOUS"]
            keep.append(arr)
            setattr(a, name, arr.ctypes.data)
    a.constitutiveRelationType = {"evp": EVP, "evp_revised": EVP_REVISED, "linear": LINEAR, "none": NONE}[
        opts.get("constitutive_relation_type", "evp")]
    a.oceanStressType = {"quadratic": 1, "linear": 2}[opts.get("ocean_stress_type", "quadratic")]
    a.useOceanStress = int(opts.get("use_ocean_stress", True))
    a.averageVariationalStrains = int(opts.get("average_variational_strain", False))
    a.useSpecialBoundariesVelocity = int(opts.get("use_special_boundaries_velocity", False))
    a.useSpecialBoundariesVelocityMasks = int(opts.get("use_special_boundaries_velocity_masks", False))
    a.elasticTimeStep = opts["elasticTimeStep"]
    a.dynamicsTimeStep = opts["dynamicsTimeStep"]
    a.dampingTimescale = opts["dampingTimescale"]
    a.numericalInertiaCoefficient = opts.get("numericalInertiaCoefficient", 0.0)
    L.orc_subcycle_velocity_solver(C.byref(a), _i(n_subcycles))
    return step


def set_num_threads(n: int) -> None:
    """OpenMP thread count used by the oracle loops that the reference marks ``!$omp parallel do``."""
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(C.c_int(int(n)))
