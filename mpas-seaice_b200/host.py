"""ctypes mirror of the C-ABI (include/evp_b200.h) for hosts written in Python (tests, bench.py).

The Fortran host binds the very same symbols through fortran/seaice_evp_b200.F90; this module is the
Python spelling of that shim and keeps the reference's names for the lifecycle
(src/shared/mpas_seaice_mesh_pool.F:76-281, src/shared/mpas_seaice_velocity_solver.F:2404-2464):

    seaice_mesh_pool_create   -> EvpSolver(...)            (evp_create)
    seaice_mesh_pool_update   -> EvpSolver.update_step     (evp_update_step)
    subcycle_velocity_solver  -> EvpSolver.run_subcycles   (evp_run_subcycles)
    seaice_mesh_pool_destroy  -> EvpSolver.destroy         (evp_destroy)

There is NO CPU fallback: if libevp_b200.so is missing or no CUDA device answers, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# EVP_B200_LIB selects another build of the same sources (e.g. the FMA-contracted perf build)
LIB_PATH = os.environ.get("EVP_B200_LIB") or os.path.join(_HERE, "csrc", "libevp_b200.so")
_lib = None

CR = {"evp": 1, "evp_revised": 2, "linear": 3, "none": 4}
OCEAN = {"quadratic": 1, "linear": 2}
SCHEME = {"variational": 1, "weak": 2}
FLAG_PIN_HOST = 1
FLAG_OVERLAP_HALO = 2
START_RESIDENT, START_FROM_REST, START_FIRST_STEP = 0, 1, 2     # evp_pre_options.cold_start

EXPORTS = (
    "evp_create", "evp_set_options", "evp_precompute_wachspress", "evp_fetch_basis", "evp_update_step",
    "evp_set_masks", "evp_run_subcycles", "evp_synchronize", "evp_fetch", "evp_destroy",
    "evp_last_error_string", "evp_comm_get_unique_id", "evp_comm_init", "evp_set_halo", "evp_last_run_ms",
    "evp_launch_count", "evp_get_stream", "evp_device_bytes", "evp_set_use_graph", "evp_profile_passes",
    "evp_host_metric_terms", "evp_set_mesh_ext", "evp_set_state", "evp_pre_subcycle", "evp_post_subcycle",
    "evp_fetch_pre", "evp_release_host_memory", "evp_set_weak_mesh", "evp_update_weak_state", "evp_fetch_weak",
    "evp_precompute_pwl", "evp_halo_mode", "evp_aggregate", "evp_fetch_aggregate", "evp_integration_rule",
)
INTEGRATION_TYPE = {"dunavant": 0, "trapezoidal": 1, "fekete": 2}     # config_wachspress_integration_type


class EvpError(RuntimeError):
    pass


class MeshDesc(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nVerticesSolve", "maxEdges",
                                        "vertexDegree")]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex",
                                             "cellVerticesAtVertex", "basisGradientU", "basisGradientV",
                                             "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric",
                                             "tanLatVertexRotatedOverRadius", "variationalDenominator",
                                             "vertexBoundaryType", "vertexBoundarySourceLocal")])


class Options(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("constitutive_relation_type", "ocean_stress_type", "use_ocean_stress",
                                        "use_special_boundaries_velocity", "device", "flags",
                                        "average_variational_strain", "strain_scheme", "stress_divergence_scheme")]
                + [(n, C.c_double) for n in ("elasticTimeStep", "dynamicsTimeStep", "dampingTimescale",
                                             "numericalInertiaCoefficient")])


STEP_FIELDS = ("solveStress", "solveVelocity", "icePressure", "uVelocity", "vVelocity", "stress11", "stress22",
               "stress12", "totalMassVertex", "totalMassVertexfVertex", "iceAreaVertex", "airStressVertexU",
               "airStressVertexV", "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "oceanStressV",
               "uOceanVelocityVertex", "vOceanVelocityVertex", "uVelocityInitial", "vVelocityInitial")
OUT_FIELDS = ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "strain11", "strain22", "strain12",
              "replacementPressure", "stressDivergenceU", "stressDivergenceV", "oceanStressCoeff")
_CELL2D = {"stress11", "stress22", "stress12", "strain11", "strain22", "strain12", "replacementPressure"}


class StepFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in STEP_FIELDS]


class OutFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUT_FIELDS]


# ---- pre-/post-subcycle on the device (include/evp_b200.h, "Widening") ----
MESH_EXT_FIELDS = ("cellsOnCell", "interiorVertex", "landIceMaskVertex", "areaCell", "areaTriangle", "fVertex")
PRE_FIELDS = ("iceAreaCellInitial", "iceAreaCell", "totalMassCell", "icePressure", "uOceanVelocity", "vOceanVelocity",
              "airStressCellU", "airStressCellV", "uAirVelocity", "vAirVelocity", "airDensity", "seaSurfaceTiltU",
              "seaSurfaceTiltV", "landIceMask", "solveStress", "solveVelocity")
_PRE_INT = {"landIceMask", "solveStress", "solveVelocity"}
POST_FIELDS = ("uVelocity", "vVelocity", "divergence", "shear", "ridgeConvergence", "ridgeShear", "principalStress1Var",
               "principalStress2Var", "oceanStressCellU", "oceanStressCellV", "oceanStressU", "oceanStressV",
               "oceanStressCoeff", "principalStress1Weak", "principalStress2Weak")
POST_FIELDS_VARIATIONAL = tuple(n for n in POST_FIELDS if not n.endswith("Weak"))
_POST_CELL = {"divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV",
              "principalStress1Weak", "principalStress2Weak"}
_POST_CELL2D = {"principalStress1Var", "principalStress2Var"}
POST_DEFAULT = ("uVelocity", "vVelocity", "divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU",
                "oceanStressCellV")       # what the model needs every step: advection, ridging, coupler
PRE_OUT_FIELDS = ("solveStress", "solveVelocity", "solveVelocityPrevious", "icePressure", "iceAreaVertex",
                  "totalMassVertex", "totalMassVertexfVertex", "airStressVertexU", "airStressVertexV",
                  "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "oceanStressV", "uOceanVelocityVertex",
                  "vOceanVelocityVertex", "uVelocityInitial", "vVelocityInitial")
_PRE_OUT_INT = {"solveStress", "solveVelocity", "solveVelocityPrevious"}
_PRE_OUT_CELL = {"solveStress", "icePressure"}


class MeshExt(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MESH_EXT_FIELDS]


class PreFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PRE_FIELDS]


class CategoryFields(C.Structure):
    _fields_ = [("nCategories", C.c_int)] + [(n, C.c_void_p) for n in ("iceAreaCategory", "iceVolumeCategory",
                                                                         "snowVolumeCategory")]


class PreOptions(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("use_air_stress", "use_surface_tilt", "geostrophic_surface_tilt",
                                       "calc_velocity_masks", "cold_start")]


class PostFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in POST_FIELDS]


class PreOutFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PRE_OUT_FIELDS]


# ---- weak operators ----
WEAK_MESH_INT = ("edgesOnCell", "verticesOnEdge", "edgesOnVertex", "cellsOnEdge")
WEAK_MESH_REAL = ("dvEdge", "dcEdge", "areaCell", "areaTriangle", "normalVectorPolygon", "normalVectorTriangle",
                  "latCellRotated", "latVertexRotated")
WEAK_FIELDS = ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain22Weak", "strain12Weak",
               "replacementPressureWeak")


class WeakMesh(C.Structure):
    _fields_ = ([("nEdges", C.c_int), ("sphere_radius", C.c_double)]
                + [(n, C.c_void_p) for n in WEAK_MESH_INT + WEAK_MESH_REAL])


class WeakFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEAK_FIELDS]


def load_library(path: str | None = None):
    """dlopen libevp_b200.so; raises EvpError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EvpError(f"{p} not found: build it with `make -C {os.path.dirname(p)}` "
                       f"(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(p)
    lib.evp_last_error_string.restype = C.c_char_p
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError if the ABI drifted
    if path is None:
        _lib = lib
    return lib


def _ptr(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], \
        f"expected contiguous {dtype} array, got {getattr(a, 'dtype', type(a))}"
    return a.ctypes.data


def integration_rule(integration_type="dunavant", order=8):
    """evp_integration_rule: (u, v, weights, normalizationFactor) of get_integration_factors (wachspress.F:1224-1287)
    as the library holds them; host-only."""
    lib = load_library()
    u, v, w = np.zeros(64), np.zeros(64), np.zeros(64)
    n, norm = C.c_int(0), C.c_double(0.0)
    rc = lib.evp_integration_rule(C.c_int(INTEGRATION_TYPE[integration_type]), C.c_int(int(order)), C.byref(n),
                                  C.c_void_p(u.ctypes.data), C.c_void_p(v.ctypes.data), C.c_void_p(w.ctypes.data),
                                  C.byref(norm))
    if rc != 0:
        raise EvpError(f"libevp_b200 error {rc}: {lib.evp_last_error_string().decode()}")
    return u[:n.value].copy(), v[:n.value].copy(), w[:n.value].copy(), norm.value


def host_metric_terms(z_rotated: np.ndarray, sphere_radius: float) -> np.ndarray:
    """evp_host_metric_terms: tan(asin(z/R))/R with the scalar libm (position-independent bits)."""
    lib = load_library()
    z = np.ascontiguousarray(z_rotated, dtype=np.float64)
    out = np.zeros_like(z)
    rc = lib.evp_host_metric_terms(C.c_int(z.size), C.c_void_p(z.ctypes.data), C.c_double(float(sphere_radius)),
                                   C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise EvpError(f"libevp_b200 error {rc}: {lib.evp_last_error_string().decode()}")
    return out


def make_options(opts: dict, device: int = -1, pin_host: bool = False) -> Options:
    o = Options()
    o.constitutive_relation_type = CR[opts.get("constitutive_relation_type", "evp")]
    o.ocean_stress_type = OCEAN[opts.get("ocean_stress_type", "quadratic")]
    o.use_ocean_stress = int(opts.get("use_ocean_stress", True))
    o.use_special_boundaries_velocity = int(opts.get("use_special_boundaries_velocity", False))
    o.device = device
    o.flags = (FLAG_PIN_HOST if pin_host else 0) | (FLAG_OVERLAP_HALO if opts.get("overlap_halo", False) else 0)
    o.average_variational_strain = int(opts.get("average_variational_strain", False))
    o.strain_scheme = SCHEME[opts.get("strain_scheme", "variational")]
    o.stress_divergence_scheme = SCHEME[opts.get("stress_divergence_scheme", "variational")]
    o.elasticTimeStep = opts["elasticTimeStep"]
    o.dynamicsTimeStep = opts["dynamicsTimeStep"]
    o.dampingTimescale = opts["dampingTimescale"]
    o.numericalInertiaCoefficient = opts.get("numericalInertiaCoefficient", 0.0)
    return o


class EvpSolver:
    """One handle <-> one GPU <-> one block.  ``mesh`` is a meshgen.Mesh (or any mapping with the
    Registry mesh fields); ``var`` holds the Registry ``velocity_variational`` static fields
    (cellVerticesAtVertex, tanLatVertexRotatedOverRadius, variationalDenominator and, unless
    ``local_coords`` is given, the five basis arrays)."""

    def __init__(self, mesh, var, opts, *, device=-1, pin_host=False, local_coords=None,
                 integration=("dunavant", 8), special_boundaries=None, n_vertices_solve=None,
                 n_cells_solve=None, basis="wachspress"):
        self.lib = load_library()
        self.nCells, self.nVertices = int(mesh["nCells"]), int(mesh["nVertices"])
        self.maxEdges, self.vertexDegree = int(mesh["maxEdges"]), int(mesh["vertexDegree"])
        md = MeshDesc()
        md.nCells = self.nCells
        md.nCellsSolve = self.nCells if n_cells_solve is None else int(n_cells_solve)
        md.nVertices = self.nVertices
        md.nVerticesSolve = self.nVertices if n_vertices_solve is None else int(n_vertices_solve)
        md.maxEdges, md.vertexDegree = self.maxEdges, self.vertexDegree
        keep = []

        def put(name, arr, dtype):
            keep.append(arr)
            setattr(md, name, _ptr(arr, dtype))

        put("nEdgesOnCell", mesh["nEdgesOnCell"], np.int32)
        put("verticesOnCell", mesh["verticesOnCell"], np.int32)
        put("cellsOnVertex", mesh["cellsOnVertex"], np.int32)
        pure_weak = opts.get("stress_divergence_scheme", "variational") == "weak" and "cellVerticesAtVertex" not in var
        if not pure_weak:
            put("cellVerticesAtVertex", var["cellVerticesAtVertex"], np.int32)
        if local_coords is None and not pure_weak:
            for n in ("basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV",
                      "basisIntegralsMetric"):
                put(n, var[n], np.float64)
        if not pure_weak:
            put("tanLatVertexRotatedOverRadius", var["tanLatVertexRotatedOverRadius"], np.float64)
            put("variationalDenominator", var["variationalDenominator"], np.float64)
        if special_boundaries is not None:
            put("vertexBoundaryType", special_boundaries[0], np.int32)
            put("vertexBoundarySourceLocal", special_boundaries[1], np.int32)
        self.opts = dict(opts)
        self._o = make_options(opts, device, pin_host)
        self._h = C.c_void_p()
        self._check(self.lib.evp_create(C.byref(self._h), C.byref(md), C.byref(self._o)))
        if local_coords is not None and basis == "pwl":
            xl, yl = local_coords
            self._check(self.lib.evp_precompute_pwl(
                self._h, C.c_void_p(_ptr(xl, np.float64)), C.c_void_p(_ptr(yl, np.float64)),
                C.c_void_p(_ptr(mesh["edgesOnCell"], np.int32)), C.c_void_p(_ptr(mesh["dvEdge"], np.float64)),
                C.c_int(int(mesh["nEdges"])), C.c_void_p(_ptr(mesh["areaCell"], np.float64))))
        elif local_coords is not None:
            xl, yl = local_coords
            itype = INTEGRATION_TYPE[integration[0]]
            self._check(self.lib.evp_precompute_wachspress(self._h, C.c_void_p(_ptr(xl, np.float64)),
                                                           C.c_void_p(_ptr(yl, np.float64)),
                                                           C.c_int(itype), C.c_int(integration[1])))

    # -- error handling ------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise EvpError(f"libevp_b200 error {rc}: {self.lib.evp_last_error_string().decode()}")

    # -- lifecycle -----------------------------------------------------------------------------
    def set_options(self, opts, pin_host=None):
        self.opts = dict(opts)
        pin = bool(self._o.flags & FLAG_PIN_HOST) if pin_host is None else pin_host
        self._o = make_options(opts, self._o.device, pin)
        self._check(self.lib.evp_set_options(self._h, C.byref(self._o)))

    def update_step(self, step):
        sf = StepFields()
        self._keep = []
        for n in STEP_FIELDS:
            a = step.get(n)
            dt = np.int32 if n in ("solveStress", "solveVelocity") else np.float64
            self._keep.append(a)
            setattr(sf, n, _ptr(a, dt))
        self._check(self.lib.evp_update_step(self._h, C.byref(sf)))

    def set_masks(self, solve_stress, solve_velocity):
        self._check(self.lib.evp_set_masks(self._h, C.c_void_p(_ptr(solve_stress, np.int32)),
                                           C.c_void_p(_ptr(solve_velocity, np.int32))))

    def run_subcycles(self, n):
        self._check(self.lib.evp_run_subcycles(self._h, C.c_int(int(n))))

    def synchronize(self):
        self._check(self.lib.evp_synchronize(self._h))

    def fetch(self, into=None, names=OUT_FIELDS):
        """Blocking copy of the outputs; allocates Registry-shaped arrays unless ``into`` has them."""
        of = OutFields()
        out = {} if into is None else into
        for n in names:
            a = out.get(n)
            if a is None:
                a = np.zeros((self.nCells + 1, self.maxEdges)) if n in _CELL2D else np.zeros(self.nVertices + 1)
                out[n] = a
            setattr(of, n, _ptr(a, np.float64))
        self._check(self.lib.evp_fetch(self._h, C.byref(of)))
        return out

    def fetch_basis(self):
        shape = (self.nCells + 1, self.maxEdges, self.maxEdges)
        arrs = [np.zeros(shape) for _ in range(5)]
        self._check(self.lib.evp_fetch_basis(self._h, *[C.c_void_p(a.ctypes.data) for a in arrs]))
        return dict(zip(("basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV",
                         "basisIntegralsMetric"), arrs))

    def last_run_ms(self):
        ms = C.c_float(0)
        self._check(self.lib.evp_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def profile_passes(self, n):
        """(cell_ms, vertex_ms, other_ms) averaged over n un-graphed subcycles (CUDA events)."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self._check(self.lib.evp_profile_passes(self._h, C.c_int(int(n)), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def launch_count(self, n):
        c = C.c_int(0)
        self._check(self.lib.evp_launch_count(self._h, C.c_int(n), C.byref(c)))
        return c.value

    def device_bytes(self):
        b = C.c_ulonglong(0)
        self._check(self.lib.evp_device_bytes(self._h, C.byref(b)))
        return b.value

    def set_use_graph(self, flag):
        self._check(self.lib.evp_set_use_graph(self._h, C.c_int(int(flag))))

    # -- pre-/post-subcycle on the device --------------------------------------------------------------
    def set_mesh_ext(self, mesh, interior_vertex, land_ice_mask_vertex=None):
        """evp_set_mesh_ext: the mesh fields velocity_solver_pre/post_subcycle read."""
        e = MeshExt()
        self._ext_keep = dict(cellsOnCell=mesh["cellsOnCell"], interiorVertex=interior_vertex,
                              landIceMaskVertex=land_ice_mask_vertex, areaCell=mesh["areaCell"],
                              areaTriangle=mesh["areaTriangle"], fVertex=mesh["fVertex"])
        for n, a in self._ext_keep.items():
            setattr(e, n, _ptr(a, np.int32 if n in ("cellsOnCell", "interiorVertex", "landIceMaskVertex") else np.float64))
        self._check(self.lib.evp_set_mesh_ext(self._h, C.byref(e)))

    def set_state(self, prev):
        """evp_set_state: seed (uVelocity, vVelocity, stress11/22/12, solveVelocityPrevious), any may be absent."""
        g = lambda n, dt: C.c_void_p(_ptr(prev.get(n), dt))
        self._check(self.lib.evp_set_state(self._h, g("uVelocity", np.float64), g("vVelocity", np.float64),
                                           g("stress11", np.float64), g("stress22", np.float64),
                                           g("stress12", np.float64), g("solveVelocityPrevious", np.int32)))

    def pre_subcycle(self, cells, *, use_air_stress=True, use_surface_tilt=True, geostrophic_surface_tilt=True,
                     calc_velocity_masks=True, cold_start=False):
        """evp_pre_subcycle: velocity_solver_pre_subcycle on the device from CELL fields.  ``cold_start``: False /
        START_RESIDENT, True / START_FROM_REST (ice at rest, not new ice) or START_FIRST_STEP (the reference's first
        step without a restart file: solveVelocityPrevious = 0, solved vertices start at the ocean velocity)."""
        pf = PreFields()
        self._pre_keep = []
        for n in PRE_FIELDS:
            a = cells.get(n)
            self._pre_keep.append(a)
            setattr(pf, n, _ptr(a, np.int32 if n in _PRE_INT else np.float64))
        po = PreOptions(int(use_air_stress), int(use_surface_tilt), int(geostrophic_surface_tilt),
                        int(calc_velocity_masks), int(cold_start))
        self._check(self.lib.evp_pre_subcycle(self._h, C.byref(pf), C.byref(po)))

    def aggregate(self, ice_area_category, ice_volume_category, snow_volume_category, hibler_strength=True):
        """evp_aggregate: aggregate_mass_and_area (and the Hibler strength, with the device's exp()) on the device from
        (nCells + 1, nCategories) category arrays; a following pre_subcycle may leave iceAreaCell, iceAreaCellInitial,
        totalMassCell and icePressure out."""
        cf = CategoryFields()
        cf.nCategories = int(ice_area_category.shape[1])
        self._agg_keep = (ice_area_category, ice_volume_category, snow_volume_category)
        for n, a in zip(("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory"), self._agg_keep):
            assert a.shape == (self.nCells + 1, cf.nCategories)
            setattr(cf, n, _ptr(a, np.float64))
        self._check(self.lib.evp_aggregate(self._h, C.byref(cf), C.c_int(int(hibler_strength))))

    def fetch_aggregate(self, ice_pressure=True):
        names = ("iceAreaCell", "iceVolumeCell", "snowVolumeCell", "totalMassCell") + (("icePressure",) if ice_pressure else ())
        out = {n: np.zeros(self.nCells + 1) for n in names}
        ptrs = [C.c_void_p(out[n].ctypes.data) for n in names] + ([] if ice_pressure else [C.c_void_p()])
        self._check(self.lib.evp_fetch_aggregate(self._h, *ptrs))
        return out

    def post_subcycle(self, into=None, names=POST_DEFAULT):
        """evp_post_subcycle: velocity_solver_post_subcycle on the device + copy of the wanted results."""
        of = PostFields()
        out = {} if into is None else into
        for n in names:
            a = out.get(n)
            if a is None:
                if n in _POST_CELL2D:
                    a = np.zeros((self.nCells + 1, self.maxEdges))
                else:
                    a = np.zeros((self.nCells if n in _POST_CELL else self.nVertices) + 1)
                out[n] = a
            setattr(of, n, _ptr(a, np.float64))
        self._check(self.lib.evp_post_subcycle(self._h, C.byref(of)))
        return out

    def fetch_pre(self, names=PRE_OUT_FIELDS):
        of = PreOutFields()
        out = {}
        for n in names:
            size = (self.nCells if n in _PRE_OUT_CELL else self.nVertices) + 1
            out[n] = np.zeros(size, dtype=np.int32 if n in _PRE_OUT_INT else np.float64)
            setattr(of, n, _ptr(out[n], out[n].dtype.type))
        self._check(self.lib.evp_fetch_pre(self._h, C.byref(of)))
        return out

    # -- weak operators ------------------------------------------------------------------------------
    def set_weak_mesh(self, mesh, weak):
        """evp_set_weak_mesh; ``weak`` holds verticesOnEdge, edgesOnVertex, normalVectorPolygon/Triangle,
        latCellRotated, latVertexRotated (weakmesh.weak_fields), the rest comes from ``mesh``."""
        wm = WeakMesh()
        wm.nEdges = int(mesh["nEdges"])
        wm.sphere_radius = float(mesh["sphere_radius"]) if mesh["on_a_sphere"] else 0.0
        self._weak_keep = []
        for n in WEAK_MESH_INT + WEAK_MESH_REAL:
            a = weak[n] if n in weak else mesh[n]
            self._weak_keep.append(a)
            setattr(wm, n, _ptr(a, np.int32 if n in WEAK_MESH_INT else np.float64))
        self._check(self.lib.evp_set_weak_mesh(self._h, C.byref(wm)))

    def update_weak_state(self, step):
        wf = WeakFields()
        for n in WEAK_FIELDS[:3]:
            setattr(wf, n, _ptr(step[n], np.float64))
        self._check(self.lib.evp_update_weak_state(self._h, C.byref(wf)))

    def fetch_weak(self):
        wf = WeakFields()
        out = {n: np.zeros(self.nCells + 1) for n in WEAK_FIELDS}
        for n in WEAK_FIELDS:
            setattr(wf, n, _ptr(out[n], np.float64))
        self._check(self.lib.evp_fetch_weak(self._h, C.byref(wf)))
        return out

    # -- multi-GPU -----------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.evp_comm_get_unique_id(buf)
        if rc != 0:
            raise EvpError(f"libevp_b200 error {rc}: {lib.evp_last_error_string().decode()}")
        return buf.raw

    def comm_init(self, rank, n_ranks, unique_id: bytes):
        self._check(self.lib.evp_comm_init(self._h, C.c_int(rank), C.c_int(n_ranks), C.c_char_p(unique_id)))

    def set_halo(self, neighbour_rank, send_offset, send_index, recv_offset, recv_index):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in
                (neighbour_rank, send_offset, send_index, recv_offset, recv_index)]
        self._check(self.lib.evp_set_halo(self._h, C.c_int(len(arrs[0])), *[C.c_void_p(a.ctypes.data) for a in arrs]))

    def halo_mode(self):
        """evp_halo_mode: 'none', 'nccl (<why not peer-to-peer>)' or 'p2p' -- the exchange the next run uses."""
        mode, why = C.c_int(0), C.create_string_buffer(256)
        self._check(self.lib.evp_halo_mode(self._h, C.byref(mode), why, C.c_int(256)))
        name = {0: "none", 1: "nccl", 2: "p2p"}[mode.value]
        return f"{name} ({why.value.decode()})" if why.value else name

    def release_host_memory(self):
        """evp_release_host_memory: undo the page-locking of every array passed under pin_host."""
        self._check(self.lib.evp_release_host_memory(self._h))

    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.evp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # reference-named aliases
    seaice_mesh_pool_update = update_step
    subcycle_velocity_solver = run_subcycles
    seaice_mesh_pool_destroy = destroy


def seaice_mesh_pool_create(mesh, var, opts, **kw) -> EvpSolver:
    return EvpSolver(mesh, var, opts, **kw)
