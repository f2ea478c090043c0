"""ctypes front-end of oracle/ir_oracle.c (incremental-remapping transport).  TEST INFRASTRUCTURE -- see the
header of that file and of oracle/__init__.py for who may import this and for the parity status."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

GEOM_NAMES = ("x", "y", "xx", "xy", "yy", "xxx", "xxy", "xyy", "yyy", "xxxx", "xxxy", "xxyy", "xyyy", "yyyy")
ERRORS = {1: "edge orientation", 2: "cell orientation", 3: "parallel basis edges", 4: "negative mass at a quadrature point",
          5: "negative mass", 6: "too many parents", 7: "too many triangles", 8: "bad argument",
          9: "tracer conservation error", 10: "monotonicity violation"}


class _GeomArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nEdges", "maxEdges", "vertexDegree",
                                        "on_a_sphere", "rotate_cartesian_grid")]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "verticesOnCell", "cellsOnEdge", "verticesOnEdge",
                                             "edgesOnVertex", "xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex",
                                             "xEdge", "yEdge", "zEdge", "dcEdge", "dvEdge",
                                             "transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "remapEdge",
                                             "cellsOnEdgeRemap", "edgesOnEdgeRemap", "xVertexOnEdge", "yVertexOnEdge",
                                             "minLengthEdgesOnVertex")]
                + [("geomAvg", C.c_void_p * 14)])


class _Tracer(C.Structure):
    _fields_ = [("nLayers", C.c_int), ("parent", C.c_int), ("nParents", C.c_int), ("hasChild", C.c_int),
                ("volumeLike", C.c_int), ("array", C.c_void_p)]


class _RunArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nEdges", "maxEdges", "vertexDegree",
                                        "nCategories", "nQuadPoints", "on_a_sphere", "rotate_cartesian_grid")]
                + [("dt", C.c_double)]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge",
                                             "verticesOnEdge", "areaCell", "dcEdge", "coeffsReconstruct",
                                             "transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "xVertexOnEdge",
                                             "yVertexOnEdge", "remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap")]
                + [("geomAvg", C.c_void_p * 14)]
                + [("uVelocity", C.c_void_p), ("vVelocity", C.c_void_p), ("nTracers", C.c_int),
                   ("tracers", C.POINTER(_Tracer))]
                + [(n, C.c_void_p) for n in ("xTriangleOut", "yTriangleOut", "triangleAreaOut", "iCellTriangleOut",
                                             "edgeFluxMassOut", "maskEdgeOut", "xGradOut", "yGradOut")]
                + [("gradTracerOut", C.c_int), ("conservationCheck", C.c_int), ("monotonicityCheck", C.c_int)]
                + [(n, C.c_void_p) for n in ("sumInitOut", "sumFinalOut", "consErrOut", "monoErrOut", "monoValOut")])


def _ptr(a, dtype):
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data


def init_geometry(mesh, irf, n_cells_solve=None, rotate=False, check=True):
    """orc_ir_init_geometry on ``mesh`` (meshgen.Mesh) + ``irf`` (mpas_seaice_b200.irmesh.ir_fields)."""
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    g = _GeomArgs()
    g.nCells, g.nVertices, g.nEdges, g.maxEdges, g.vertexDegree = nC, nV, nE, M, mesh.vertexDegree
    g.nCellsSolve = nC if n_cells_solve is None else int(n_cells_solve)
    g.on_a_sphere, g.rotate_cartesian_grid = int(bool(mesh.on_a_sphere)), int(bool(rotate))
    keep = []
    for name in ("nEdgesOnCell", "edgesOnCell", "verticesOnCell", "cellsOnEdge"):
        setattr(g, name, _ptr(mesh[name], np.int32))
    for name in ("verticesOnEdge", "edgesOnVertex"):
        setattr(g, name, _ptr(irf[name], np.int32))
    for name in ("xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex", "dcEdge", "dvEdge"):
        setattr(g, name, _ptr(mesh[name], np.float64))
    for name in ("xEdge", "yEdge", "zEdge"):
        setattr(g, name, _ptr(irf[name], np.float64))
    out = dict(transGlobalToCell=np.zeros((max(nC, 1), 3, 3)),
               xVertexOnCell=np.zeros((nC + 1, M)), yVertexOnCell=np.zeros((nC + 1, M)),
               remapEdge=np.zeros(nE + 1, np.int32),
               cellsOnEdgeRemap=np.zeros((nE + 1, 6), np.int32), edgesOnEdgeRemap=np.zeros((nE + 1, 6), np.int32),
               xVertexOnEdge=np.zeros((nE + 1, 8)), yVertexOnEdge=np.zeros((nE + 1, 8)),
               minLengthEdgesOnVertex=np.zeros(nV + 1))
    for name, arr in out.items():
        setattr(g, name, arr.ctypes.data)
    geom = {n: np.zeros(nC + 1) for n in GEOM_NAMES}
    for k, n in enumerate(GEOM_NAMES):
        g.geomAvg[k] = geom[n].ctypes.data
    keep.append(geom)
    L = lib()
    L.orc_ir_init_geometry.restype = C.c_int
    err = L.orc_ir_init_geometry(C.byref(g))
    if check and err:
        raise RuntimeError("orc_ir_init_geometry: " + ERRORS.get(err, str(err)))
    out["geomAvg"] = geom
    out["error"] = err
    return out


class Tracer:
    """One entry of the reference's tracer linked list (incremental_remap_tracers.F:26-110): ``array`` is
    (nCells+1, nCategories, nLayers); parent = index of the parent in the list (None for the mass-like field)."""

    def __init__(self, name, array, parent=None, volume_like=False):
        assert array.ndim == 3 and array.dtype == np.float64 and array.flags["C_CONTIGUOUS"]
        self.name, self.array, self.parent, self.volume_like = name, array, parent, volume_like


def default_tracers(n_cells, n_categories, n_ice_layers=0, n_snow_layers=0, rng=None):
    """The always-present tracers of seaice_add_tracers_to_linked_list (incremental_remap_tracers.F:196-202),
    zero-filled (or random positive with ``rng``); the enthalpy / salinity layers only when asked for."""
    def new(nl):
        a = np.zeros((n_cells + 1, n_categories, nl))
        if rng is not None:
            a[:n_cells] = rng.uniform(0.1, 1.0, size=(n_cells, n_categories, nl))
        return a
    tr = [Tracer("iceAreaCategory", new(1)),
          Tracer("iceVolumeCategory", new(1), 0, True),
          Tracer("snowVolumeCategory", new(1), 0, True),
          Tracer("surfaceTemperature", new(1), 0)]
    if n_ice_layers:
        tr += [Tracer("iceEnthalpy", new(n_ice_layers), 1), Tracer("iceSalinity", new(n_ice_layers), 1)]
    if n_snow_layers:
        tr += [Tracer("snowEnthalpy", new(n_snow_layers), 2)]
    return tr


def run(mesh, irf, geom, tracers, u, v, dt, n_quad_points=6, n_cells_solve=None, rotate=False, diagnostics=False,
        grad_tracer=0, check=True, conservation_check=0, monotonicity_check=0):
    """One call of seaice_run_advection_incremental_remap (single block, no halo update) IN PLACE on the tracers.
    ``conservation_check`` (1: sums and check, 2: sums only) / ``monotonicity_check`` (1: the reference's in-place
    extension of the bounds, 2: the order-independent one) switch on config_conservation_check /
    config_monotonicity_check; the returned dict then holds ``sumInit`` / ``sumFinal`` (one (nCategories, nLayers) array
    per tracer), ``consErr`` [violated, tracer, iCat, iLayer], ``monoErr`` [0 / 1 min / 2 max, tracer, iLayer, iCat,
    iCell] and ``monoVal`` [new value, bound, tolerance]."""
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    nK = tracers[0].array.shape[1]
    a = _RunArgs()
    a.nCells, a.nVertices, a.nEdges, a.maxEdges, a.vertexDegree = nC, nV, nE, M, mesh.vertexDegree
    a.nCellsSolve = nC if n_cells_solve is None else int(n_cells_solve)
    a.nCategories, a.nQuadPoints = nK, n_quad_points
    a.on_a_sphere, a.rotate_cartesian_grid = int(bool(mesh.on_a_sphere)), int(bool(rotate))
    a.dt = float(dt)
    for name in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge"):
        setattr(a, name, _ptr(mesh[name], np.int32))
    a.verticesOnEdge = _ptr(irf["verticesOnEdge"], np.int32)
    a.areaCell, a.dcEdge = _ptr(mesh.areaCell, np.float64), _ptr(mesh.dcEdge, np.float64)
    a.coeffsReconstruct = _ptr(irf["coeffs_reconstruct"], np.float64)
    for name in ("transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge"):
        setattr(a, name, _ptr(geom[name], np.float64))
    for name in ("remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap"):
        setattr(a, name, _ptr(geom[name], np.int32))
    for k, n in enumerate(GEOM_NAMES):
        a.geomAvg[k] = _ptr(geom["geomAvg"][n], np.float64)
    assert u.shape == (nV + 1,) and v.shape == (nV + 1,)
    a.uVelocity, a.vVelocity = _ptr(u, np.float64), _ptr(v, np.float64)
    table = (_Tracer * len(tracers))()
    has_child = [False] * len(tracers)
    for t in tracers:
        if t.parent is not None:
            has_child[t.parent] = True
    n_parents = []
    for i, t in enumerate(tracers):
        assert t.array.shape[:2] == (nC + 1, nK)
        n_parents.append(0 if t.parent is None else n_parents[t.parent] + 1)
        table[i].nLayers = t.array.shape[2]
        table[i].parent = -1 if t.parent is None else t.parent
        table[i].nParents = n_parents[i]
        table[i].hasChild = int(has_child[i])
        table[i].volumeLike = int(t.volume_like)
        table[i].array = t.array.ctypes.data
    a.nTracers = len(tracers)
    a.tracers = table
    diag = {}
    a.gradTracerOut = -1
    if diagnostics:
        nL0 = tracers[0].array.shape[2]
        diag = dict(xTriangle=np.zeros((nE, 6, n_quad_points)), yTriangle=np.zeros((nE, 6, n_quad_points)),
                    triangleArea=np.zeros((nE, 6)), iCellTriangle=np.zeros((nE, 6), np.int32),
                    edgeFluxMass=np.zeros((nE, nK, nL0)), maskEdge=np.zeros(nE, np.int32),
                    xGrad=np.zeros_like(tracers[grad_tracer].array), yGrad=np.zeros_like(tracers[grad_tracer].array))
        a.xTriangleOut, a.yTriangleOut = diag["xTriangle"].ctypes.data, diag["yTriangle"].ctypes.data
        a.triangleAreaOut, a.iCellTriangleOut = diag["triangleArea"].ctypes.data, diag["iCellTriangle"].ctypes.data
        a.edgeFluxMassOut, a.maskEdgeOut = diag["edgeFluxMass"].ctypes.data, diag["maskEdge"].ctypes.data
        a.xGradOut, a.yGradOut = diag["xGrad"].ctypes.data, diag["yGrad"].ctypes.data
        a.gradTracerOut = grad_tracer
    a.conservationCheck, a.monotonicityCheck = int(conservation_check), int(monotonicity_check)
    if conservation_check or monotonicity_check:
        n_sum = sum(nK * t.array.shape[2] for t in tracers)
        flat_i, flat_f = np.zeros(n_sum), np.zeros(n_sum)
        diag["consErr"], diag["monoErr"], diag["monoVal"] = np.zeros(4, np.int32), np.zeros(5, np.int32), np.zeros(3)
        a.sumInitOut, a.sumFinalOut = flat_i.ctypes.data, flat_f.ctypes.data
        a.consErrOut, a.monoErrOut, a.monoValOut = (diag["consErr"].ctypes.data, diag["monoErr"].ctypes.data,
                                                    diag["monoVal"].ctypes.data)
    L = lib()
    L.orc_ir_run.restype = C.c_int
    err = L.orc_ir_run(C.byref(a))
    if conservation_check or monotonicity_check:
        off, diag["sumInit"], diag["sumFinal"] = 0, [], []
        for t in tracers:
            n = nK * t.array.shape[2]
            diag["sumInit"].append(flat_i[off:off + n].reshape(nK, t.array.shape[2]).copy())
            diag["sumFinal"].append(flat_f[off:off + n].reshape(nK, t.array.shape[2]).copy())
            off += n
    if check and err:
        raise RuntimeError("orc_ir_run: " + ERRORS.get(err, str(err)))
    diag["error"] = err
    return diag
