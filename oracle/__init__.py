"""ctypes front-end of the CPU oracle (oracle/evp_oracle.c, oracle/evp_precompute_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; the product never does.

Parity status: the reference Fortran cannot be built here and stores no outputs; the EVP oracle is pinned by
outputs of the reference's own source executed by an interpreter (tests/golden/fortran_subset.py,
tests/test_golden.py, tests/test_refexec_init.py, tests/test_refexec_step.py: bit for bit) and by the reference's
analytic known answers (tests/test_oracle_kat.py).  See the headers of the C files.
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
    srcs = [os.path.join(_HERE, f) for f in ("evp_oracle.c", "evp_precompute_oracle.c", "ir_oracle.c", "upwind_oracle.c", "Makefile",
                                              "quadrature_tables.inc")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def build_variant(name: str = "o3native") -> str:
    """A second build of the same sources for TIMING only ("a second -O3 -march=native build
    is reported separately and labelled as such"): gcc -O3 -march=native with the compiler's default contraction, i.e.
    what a production CPU build would do.  Compiled where it runs (-march=native is only valid for the machine that
    compiles it), into oracle/_build/; never used as the checker."""
    import platform
    import hashlib
    if name != "o3native":
        raise ValueError(name)
    tag = hashlib.md5((platform.node() + platform.processor() + platform.machine()).encode()).hexdigest()[:8]
    out = os.path.join(_HERE, "_build", f"liboracle_{name}_{tag}.so")
    srcs = [os.path.join(_HERE, f) for f in ("evp_oracle.c", "evp_precompute_oracle.c", "ir_oracle.c", "upwind_oracle.c")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-march=native", "-fPIC", "-fopenmp", "-shared", "-o", out, *srcs, "-lm"], check=True)
    return out


def load_variant(name: str = "o3native"):
    return C.CDLL(build_variant(name))


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
        itype = {"dunavant": 0, "trapezoidal": 1, "fekete": 2}[integration_type]
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
    err = L.orc_get_integration_factors(_i({"dunavant": 0, "trapezoidal": 1, "fekete": 2}[integration_type]), _i(order),
                                        C.byref(n), _p(u), _p(v), _p(w), C.byref(norm))
    assert err == 0
    return u[:n.value].copy(), v[:n.value].copy(), w[:n.value].copy(), norm.value


# ---------------------------------------------------------------------------------------------
# subcycle
# ---------------------------------------------------------------------------------------------

_WEAK_STATIC = ("edgesOnCell", "verticesOnEdge", "edgesOnVertex", "cellsOnEdge", "dvEdge", "dcEdge", "areaTriangle",
                "normalVectorPolygon", "normalVectorTriangle", "latCellRotated", "latVertexRotated")
_WEAK_STEP = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak",
              "replacementPressureWeak", "strain11Vertex", "strain22Vertex", "strain12Vertex")


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
        + [("strainScheme", C.c_int), ("stressDivergenceScheme", C.c_int), ("sphere_radius", C.c_double)]
        + [(n, C.c_void_p) for n in _WEAK_STATIC + _WEAK_STEP]
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


def subcycle_velocity_solver(mesh, var, step, opts, n_subcycles, library=None):
    """subcycle_velocity_solver (velocity_solver.F:2404-2464) on host arrays, IN PLACE on ``step``.

    mesh: meshgen.Mesh; var: dict from init_variational; step: dict of per-step fields (see
    mpas_seaice_b200.synthetic.pre_subcycle); opts: dict with constitutive_relation_type ('evp' |
    'evp_revised' | 'linear' | 'none'), ocean_stress_type, use_ocean_stress, average_variational_strain,
    elasticTimeStep, dynamicsTimeStep, dampingTimescale, numericalInertiaCoefficient.
    """
    L = library if library is not None else lib()      # `library`: a timing variant (load_variant), never the checker
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
    # weak operators: opts strain_scheme / stress_divergence_scheme = 'weak'; the static fields come from
    # var["weak"] (mpas_seaice_b200.weakmesh.weak_fields) or the mesh, the weak state from ``step``
    scheme = {"variational": 1, "weak": 2}
    a.strainScheme = scheme[opts.get("strain_scheme", "variational")]
    a.stressDivergenceScheme = scheme[opts.get("stress_divergence_scheme", "variational")]
    a.sphere_radius = 0.0
    if a.strainScheme == 2:
        a.sphere_radius = float(mesh.sphere_radius) if mesh.on_a_sphere else 0.0
        weak = var["weak"]
        for name in _WEAK_STATIC:
            arr = weak[name] if name in weak else mesh[name]
            assert arr.flags["C_CONTIGUOUS"]
            keep.append(arr)
            setattr(a, name, arr.ctypes.data)
        nC, nV = mesh.nCells, mesh.nVertices
        for name in _WEAK_STEP:
            if name not in step:
                step[name] = np.zeros((nV if name.endswith("Vertex") else nC) + 1)
            keep.append(step[name])
            setattr(a, name, step[name].ctypes.data)
    L.orc_subcycle_velocity_solver(C.byref(a), _i(n_subcycles))
    return step


def set_num_threads(n: int) -> None:
    """OpenMP thread count used by the oracle loops that the reference marks ``!$omp parallel do``."""
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(C.c_int(int(n)))


# ---------------------------------------------------------------------------------------------
# pre- / post-subcycle (the host side of the boundary, restated so tests can build identical inputs)
# ---------------------------------------------------------------------------------------------

def boundary_source_local(index_to_id, boundary_type, boundary_source, out=None):
    """init_special_boundaries_velocity / _tracers (special_boundaries.F:83-150, 164-250); arrays with the extra slot"""
    n = len(index_to_id) - 1
    res = np.zeros(n + 1, dtype=np.int32) if out is None else out
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (index_to_id, boundary_type, boundary_source)]
    L = lib()
    L.orc_boundary_source_local.restype = C.c_int
    if L.orc_boundary_source_local(_i(n), _p(a[0]), _p(a[1]), _p(a[2]), _p(res)) != 0:
        raise ValueError("global IDs outside 1..n: the reference's globalToLocalID table would be indexed out of bounds")
    return res


def set_special_boundaries_tracers(boundary_type, source_local, ice_area_category, ice_volume_category, snow_volume_category):
    """seaice_set_special_boundaries_tracers (special_boundaries.F:415-485), IN PLACE on (nCells+1, nCategories, 1) arrays"""
    nC = len(boundary_type) - 1
    n = int(np.prod(ice_area_category.shape[1:]))
    for a in (ice_area_category, ice_volume_category, snow_volume_category):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape[0] == nC + 1
    lib().orc_set_special_boundaries_tracers(_i(nC), _i(n), _p(np.ascontiguousarray(boundary_type, dtype=np.int32)),
                                             _p(np.ascontiguousarray(source_local, dtype=np.int32)),
                                             _p(ice_area_category), _p(ice_volume_category), _p(snow_volume_category))


def land_ice_mask_vertex(mesh, land_ice_mask, n_vertices_solve=None):
    """init_ice_shelve_vertex_mask (velocity_solver.F:481-544): (nVertices+1) int32"""
    nV = mesh.nVertices
    out = np.zeros(nV + 1, dtype=np.int32)
    land = np.ascontiguousarray(land_ice_mask, dtype=np.int32)
    assert land.shape == (mesh.nCells + 1,)
    lib().orc_ice_shelve_vertex_mask(_i(nV), _i(nV if n_vertices_solve is None else n_vertices_solve), _i(mesh.vertexDegree),
                                     _p(mesh.cellsOnVertex), _p(land), _p(out))
    return out


def dynamically_locked_cells_mask(mesh, interior_vertex):
    """dynamically_locked_cell_mask (velocity_solver.F:402-467): (nCells+1) int32"""
    nC = mesh.nCells
    out = np.zeros(nC + 1, dtype=np.int32)
    iv = np.ascontiguousarray(interior_vertex, dtype=np.int32)
    assert iv.shape == (mesh.nVertices + 1,)
    lib().orc_dynamically_locked_cell_mask(_i(nC), _i(mesh.maxEdges), _p(mesh.nEdgesOnCell), _p(mesh.verticesOnCell),
                                           _p(iv), _p(out))
    return out


def pre_subcycle(mesh, state, config_dt, *, n_elastic=120, use_air_stress=True, use_ocean_stress=True,
                 use_surface_tilt=True, interior_vertex=None, prev=None, geostrophic_surface_tilt=True,
                 land_ice_mask=None, land_ice_mask_vertex=None):
    """velocity_solver_pre_subcycle (velocity_solver.F:613-671) from a cold start, single category,
    Hibler strength, constant_air_stress, geostrophic tilt -- call order of the reference.
    ``prev`` (uVelocity, vVelocity, stress11/22/12, solveVelocityPrevious) = the state carried from the previous
    dynamics step instead of a cold start; ``geostrophic_surface_tilt=False`` uses state["seaSurfaceTiltU/V"]
    (surface_tilt_ssh_gradient, :2024-2170).  ``land_ice_mask`` (nCells+1) / ``land_ice_mask_vertex`` (nVertices+1): the
    ocean_coupling pool's ice-shelf masks read by the calculation masks (:1023, :1131); None = no land ice.
    Returns the same dict of per-step fields as mpas_seaice_b200.synthetic.pre_subcycle."""
    L = lib()
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    z = lambda n: np.zeros(n)
    f = {}
    area = np.ascontiguousarray(state["iceAreaCell"], dtype=np.float64)
    vol = np.ascontiguousarray(state["iceVolumeCell"], dtype=np.float64)
    snow = np.ascontiguousarray(state["snowVolumeCell"], dtype=np.float64)
    mass = z(nC + 1)
    L.orc_total_mass(_i(nC + 1), _p(vol), _p(snow), _p(mass))

    def c2v(cell):
        out = z(nV + 1)
        L.orc_interpolate_cell_to_vertex(_i(nV), _i(D), _p(mesh.cellsOnVertex), _p(mesh.areaCell),
                                         _p(np.ascontiguousarray(cell, dtype=np.float64)), _p(out))
        return out

    with np.errstate(all="ignore"):
        f["iceAreaVertex"] = c2v(area)
        f["totalMassVertex"] = c2v(mass)
        land = (np.zeros(nC + 1, dtype=np.int32) if land_ice_mask is None
                else np.ascontiguousarray(land_ice_mask, dtype=np.int32))
        landv = (np.zeros(nV + 1, dtype=np.int32) if land_ice_mask_vertex is None
                 else np.ascontiguousarray(land_ice_mask_vertex, dtype=np.int32))
        assert land.shape == (nC + 1,) and landv.shape == (nV + 1,)
        ss = np.zeros(nC + 1, dtype=np.int32)
        L.orc_stress_calculation_mask(_i(nC), _i(M), _p(mesh.nEdgesOnCell), _p(mesh.cellsOnCell), _p(area), _p(mass),
                                      _p(land), _p(ss))
        if interior_vertex is None:
            interior_vertex = np.zeros(nV + 1, dtype=np.int32)
            L.orc_interior_vertices(_p(interior_vertex), _i(nV), _i(D), _i(nC), _p(mesh.cellsOnVertex))
        sv = np.zeros(nV + 1, dtype=np.int32)
        L.orc_velocity_calculation_mask(_i(nV), _i(nV), _p(interior_vertex), _p(landv), _p(f["iceAreaVertex"]),
                                        _p(f["totalMassVertex"]), _p(sv))
        f["solveStress"], f["solveVelocity"] = ss, sv
        f["uOceanVelocityVertex"] = c2v(state["uOceanVelocity"])
        f["vOceanVelocityVertex"] = c2v(state["vOceanVelocity"])
        if prev is None:
            u, v = z(nV + 1), z(nV + 1)
            svp = sv.copy()
        else:
            u = np.array(prev["uVelocity"], dtype=np.float64)
            v = np.array(prev["vVelocity"], dtype=np.float64)
            svp = np.array(prev["solveVelocityPrevious"], dtype=np.int32)
        sdu, sdv, osu, osv = z(nV + 1), z(nV + 1), z(nV + 1), z(nV + 1)
        ui, vi = z(nV + 1), z(nV + 1)
        L.orc_new_ice_velocities(_i(nV), _i(nV), _p(sv), _p(svp), _p(f["uOceanVelocityVertex"]),
                                 _p(f["vOceanVelocityVertex"]), _p(u), _p(v), _p(sdu), _p(sdv), _p(osu), _p(osv),
                                 _p(ui), _p(vi))
        f["solveVelocityPrevious"], f["uVelocityInitial"], f["vVelocityInitial"] = svp, ui, vi
        P = z(nC + 1)
        L.orc_ice_strength_hibler(_i(nC), _p(ss), _p(vol), _p(area), _p(P))
        f["icePressure"] = P
        au, av = z(nC + 1), z(nC + 1)
        if use_air_stress:
            L.orc_constant_air_stress(_i(nC + 1), _p(np.ascontiguousarray(state["uAirVelocity"])),
                                      _p(np.ascontiguousarray(state["vAirVelocity"])),
                                      _p(np.ascontiguousarray(state["airDensity"])), _p(area), _p(au), _p(av))
        f["airStressVertexU"], f["airStressVertexV"] = c2v(au), c2v(av)
        mf = z(nV + 1)
        L.orc_coriolis_force_coefficient(_i(nV + 1), _p(f["totalMassVertex"]), _p(mesh.fVertex), _p(mf))
        f["totalMassVertexfVertex"] = mf
        L.orc_ocean_stress(_i(nV), _i(nV), _i(use_ocean_stress), _p(sv), _p(f["uOceanVelocityVertex"]),
                           _p(f["vOceanVelocityVertex"]), _p(mesh.fVertex), _p(osu), _p(osv))
        f["oceanStressU"], f["oceanStressV"] = osu, osv
        tu, tv = z(nV + 1), z(nV + 1)
        if use_surface_tilt and not geostrophic_surface_tilt:
            f["seaSurfaceTiltVertexU"] = c2v(state["seaSurfaceTiltU"])
            f["seaSurfaceTiltVertexV"] = c2v(state["seaSurfaceTiltV"])
            L.orc_surface_tilt_ssh_gradient(_i(nV), _p(sv), _p(f["totalMassVertex"]), _p(f["seaSurfaceTiltVertexU"]),
                                            _p(f["seaSurfaceTiltVertexV"]), _p(tu), _p(tv))
        else:
            L.orc_surface_tilt(_i(nV), _i(nV), _i(use_surface_tilt), _p(sv), _p(mesh.fVertex), _p(f["totalMassVertex"]),
                               _p(f["uOceanVelocityVertex"]), _p(f["vOceanVelocityVertex"]), _p(tu), _p(tv))
        f["surfaceTiltForceU"], f["surfaceTiltForceV"] = tu, tv
    oc = z(nV + 1)
    e = [np.zeros((nC + 1, M)) for _ in range(3)]
    if prev is None:
        s = [np.zeros((nC + 1, M)) for _ in range(3)]
    else:
        s = [np.array(prev[k], dtype=np.float64) for k in ("stress11", "stress22", "stress12")]
    L.orc_init_subcycle_variables(_i(nC), _i(nV), _i(nV), _i(M), _p(ss), _p(sv), _p(sdu), _p(sdv), _p(u), _p(v), _p(oc),
                                  _p(e[0]), _p(e[1]), _p(e[2]), _p(s[0]), _p(s[1]), _p(s[2]))
    f.update(stressDivergenceU=sdu, stressDivergenceV=sdv, oceanStressCoeff=oc, uVelocity=u, vVelocity=v,
             strain11=e[0], strain22=e[1], strain12=e[2], stress11=s[0], stress22=s[1], stress12=s[2],
             replacementPressure=np.zeros((nC + 1, M)))
    return f


def final_divergence_shear(mesh, step):
    """seaice_final_divergence_shear_variational (variational.F:1198-1330)"""
    L = lib()
    nC = mesh.nCells
    out = [np.zeros(nC + 1) for _ in range(4)]
    L.orc_final_divergence_shear_variational(_i(nC), _i(mesh.maxEdges), _p(mesh.nEdgesOnCell), _p(step["solveStress"]),
                                             _p(step["strain11"]), _p(step["strain22"]), _p(step["strain12"]),
                                             *[_p(a) for a in out])
    return dict(zip(("divergence", "shear", "ridgeConvergence", "ridgeShear"), out))


def aggregate_mass_and_area(ice_area_category, ice_volume_category, snow_volume_category):
    """aggregate_mass_and_area (velocity_solver.F:685-752) on (nCells + 1, nCategories) arrays; returns
    iceAreaCell, iceVolumeCell, snowVolumeCell, totalMassCell."""
    a = np.ascontiguousarray(ice_area_category, dtype=np.float64)
    n, k = a.shape
    vi = np.ascontiguousarray(ice_volume_category, dtype=np.float64)
    vs = np.ascontiguousarray(snow_volume_category, dtype=np.float64)
    out = [np.zeros(n) for _ in range(4)]
    lib().orc_aggregate_mass_and_area(_i(n), _i(k), _p(a), _p(vi), _p(vs), *[_p(o) for o in out])
    return out


def hibler_strength_unmasked(state, n_cells):
    """seaiceIceStrengthConstantHiblerP * iceVolumeCell * exp(-C*(1-iceAreaCell)) for EVERY cell (libm exp), i.e.
    ice_strength (velocity_solver.F:1419-1436) before its solveStress mask: what a host hands to evp_pre_subcycle."""
    L = lib()
    ones = np.ones(n_cells + 1, dtype=np.int32)
    P = np.zeros(n_cells + 1)
    L.orc_ice_strength_hibler(_i(n_cells), _p(ones), _p(np.ascontiguousarray(state["iceVolumeCell"], dtype=np.float64)),
                              _p(np.ascontiguousarray(state["iceAreaCell"], dtype=np.float64)), _p(P))
    return P


def ocean_stress_final(mesh, step, opts, interior_vertex, n_vertices_solve=None, n_cells_solve=None):
    """ocean_stress_final (velocity_solver.F:3624-3848) on the fields of ``step`` AFTER the subcycle: refreshes
    oceanStressCoeff with the final velocities, returns (oceanStressU, oceanStressV, oceanStressCellU,
    oceanStressCellV, oceanStressCoeff)."""
    L = lib()
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    nVs = nV if n_vertices_solve is None else n_vertices_solve
    nCs = nC if n_cells_solve is None else n_cells_solve
    use_ocean = int(opts.get("use_ocean_stress", True))
    otype = {"quadratic": 1, "linear": 2}[opts.get("ocean_stress_type", "quadratic")]
    coef = np.array(step["oceanStressCoeff"], dtype=np.float64)
    L.orc_ocean_stress_coefficient(_i(nVs), _i(nV), _i(use_ocean), _i(otype), _p(step["solveVelocity"]),
                                   _p(step["iceAreaVertex"]), _p(step["uOceanVelocityVertex"]),
                                   _p(step["vOceanVelocityVertex"]), _p(step["uVelocity"]), _p(step["vVelocity"]), _p(coef))
    osu, osv = np.zeros(nV + 1), np.zeros(nV + 1)
    ocu, ocv = np.zeros(nC + 1), np.zeros(nC + 1)
    L.orc_ocean_stress_final(_i(nVs), _i(nV), _i(nCs), _i(nC), _i(M), _i(use_ocean), _p(step["solveVelocity"]), _p(coef),
                             _p(step["uOceanVelocityVertex"]), _p(step["vOceanVelocityVertex"]), _p(step["uVelocity"]),
                             _p(step["vVelocity"]), _p(mesh.fVertex), _p(step["iceAreaVertex"]), _p(mesh.nEdgesOnCell),
                             _p(mesh.verticesOnCell), _p(mesh.areaTriangle), _p(interior_vertex), _p(osu), _p(osv),
                             _p(ocu), _p(ocv))
    return osu, osv, ocu, ocv, coef


def principal_stresses(mesh, step, n_cells_solve=None):
    """principal_stresses_driver, variational branch (velocity_solver.F:3443-3610)"""
    L = lib()
    nC, M = mesh.nCells, mesh.maxEdges
    p1, p2 = np.zeros((nC + 1, M)), np.zeros((nC + 1, M))
    L.orc_principal_stresses_variational(_i(nC if n_cells_solve is None else n_cells_solve), _i(M),
                                         _p(mesh.nEdgesOnCell), _p(step["stress11"]), _p(step["stress22"]),
                                         _p(step["stress12"]), _p(step["replacementPressure"]), _p(p1), _p(p2))
    return p1, p2
