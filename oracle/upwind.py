"""ctypes front-end of oracle/upwind_oracle.c (seaice_normal_vectors and the upwind transport).  TEST INFRASTRUCTURE --
see the header of that file and of oracle/__init__.py for who may import this and for the parity status."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib


class _NormalsArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nVertices", "nVerticesSolve", "nEdges", "maxEdges", "vertexDegree",
                                        "on_a_sphere", "rotate_cartesian_grid", "removeMetricTerms")]
                + [("sphere_radius", C.c_double)]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "verticesOnEdge", "cellsOnEdge", "edgesOnVertex",
                                             "interiorVertex", "xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex",
                                             "xEdge", "yEdge", "zEdge", "normalVectorPolygon", "normalVectorTriangle",
                                             "latCellRotated", "latVertexRotated")])


class _UpwindVar(C.Structure):
    _fields_ = [("parent", C.c_int), ("childMinimum", C.c_double), ("volumeLike", C.c_int), ("array", C.c_void_p),
                ("edgeFluxOut", C.c_void_p)]


class _UpwindArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nEdges", "maxEdges", "nCategories")]
                + [("dt", C.c_double)]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "cellsOnEdge", "verticesOnEdge",
                                             "interiorEdge", "areaCell", "dvEdge", "normalVectorEdge", "uVelocity",
                                             "vVelocity")]
                + [("nVars", C.c_int), ("vars", C.POINTER(_UpwindVar)), ("edgeVelocityOut", C.c_void_p)])


def _ptr(a, dtype):
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data


def normal_vectors(mesh, edges, interior_vertex, rotate=True, remove_metric_terms=True, triangles=True,
                   n_vertices_solve=None):
    """seaice_normal_vectors (mesh.F:703) on ``mesh`` (meshgen.Mesh) with the edge arrays of ``edges`` (verticesOnEdge,
    edgesOnVertex, xEdge, yEdge, zEdge).  Returns dict(normalVectorPolygon (nCells+1, maxEdges, 2), normalVectorTriangle
    (nVertices+1, vertexDegree, 2), latCellRotated, latVertexRotated); ``triangles=False`` is
    seaice_normal_vectors_polygon alone (what the upwind transport calls, advection_upwind.F:122)."""
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    a = _NormalsArgs()
    a.nCells, a.nVertices, a.nEdges, a.maxEdges, a.vertexDegree = nC, nV, nE, M, D
    a.nVerticesSolve = nV if n_vertices_solve is None else int(n_vertices_solve)
    a.on_a_sphere = int(bool(mesh.on_a_sphere))
    a.rotate_cartesian_grid, a.removeMetricTerms = int(bool(rotate)), int(bool(remove_metric_terms))
    a.sphere_radius = float(getattr(mesh, "sphere_radius", 1.0) or 1.0)
    for name in ("nEdgesOnCell", "edgesOnCell", "cellsOnEdge"):
        setattr(a, name, _ptr(mesh[name], np.int32))
    for name in ("verticesOnEdge", "edgesOnVertex"):
        setattr(a, name, _ptr(edges[name], np.int32))
    iv = np.ascontiguousarray(interior_vertex, dtype=np.int32)
    a.interiorVertex = iv.ctypes.data
    for name in ("xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex"):
        setattr(a, name, _ptr(mesh[name], np.float64))
    for name in ("xEdge", "yEdge", "zEdge"):
        setattr(a, name, _ptr(edges[name], np.float64))
    out = dict(normalVectorPolygon=np.zeros((nC + 1, M, 2)), latCellRotated=np.zeros(nC + 1))
    a.normalVectorPolygon, a.latCellRotated = out["normalVectorPolygon"].ctypes.data, out["latCellRotated"].ctypes.data
    if triangles:
        out.update(normalVectorTriangle=np.zeros((nV + 1, D, 2)), latVertexRotated=np.zeros(nV + 1))
        a.normalVectorTriangle, a.latVertexRotated = out["normalVectorTriangle"].ctypes.data, out["latVertexRotated"].ctypes.data
    L = lib()
    L.orc_normal_vectors.restype = C.c_int
    err = L.orc_normal_vectors(C.byref(a))
    assert err == 0, err
    return out


class Var:
    """One row of the reference's tracerConnectivities table (advection_upwind.F:37-56): ``array`` (nCells+1,
    nCategories), ``parent`` = index in the table or None for 'none'."""

    def __init__(self, name, array, parent=None, volume_like=False, child_minimum=0.0):
        assert array.ndim == 2 and array.dtype == np.float64 and array.flags["C_CONTIGUOUS"]
        self.name, self.array, self.parent, self.volume_like, self.child_minimum = name, array, parent, volume_like, child_minimum


def run(mesh, vertices_on_edge, interior_edge, normal_vector_edge, variables, u, v, dt, n_cells_solve=None,
        diagnostics=False):
    """One call of seaice_run_advection_upwind (one block, halo exchange left out) IN PLACE on the variables' arrays."""
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    nK = variables[0].array.shape[1]
    a = _UpwindArgs()
    a.nCells, a.nVertices, a.nEdges, a.maxEdges, a.nCategories = nC, nV, nE, M, nK
    a.nCellsSolve = nC if n_cells_solve is None else int(n_cells_solve)
    a.dt = float(dt)
    for name in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "cellsOnEdge"):
        setattr(a, name, _ptr(mesh[name], np.int32))
    a.verticesOnEdge, a.interiorEdge = _ptr(vertices_on_edge, np.int32), _ptr(interior_edge, np.int32)
    a.areaCell, a.dvEdge = _ptr(mesh.areaCell, np.float64), _ptr(mesh.dvEdge, np.float64)
    assert normal_vector_edge.shape == (nC + 1, M, 2)
    a.normalVectorEdge = _ptr(normal_vector_edge, np.float64)
    assert u.shape == (nV + 1,) and v.shape == (nV + 1,)
    a.uVelocity, a.vVelocity = _ptr(u, np.float64), _ptr(v, np.float64)
    table = (_UpwindVar * len(variables))()
    out = {}
    if diagnostics:
        out["edgeFlux"] = [np.zeros((nE + 1, nK)) for _ in variables]
        out["edgeVelocity"] = np.zeros(nE + 1)
        a.edgeVelocityOut = out["edgeVelocity"].ctypes.data
    for i, var in enumerate(variables):
        assert var.array.shape == (nC + 1, nK)
        table[i].parent = -1 if var.parent is None else int(var.parent)
        table[i].childMinimum = float(var.child_minimum)
        table[i].volumeLike = int(bool(var.volume_like))
        table[i].array = var.array.ctypes.data
        table[i].edgeFluxOut = out["edgeFlux"][i].ctypes.data if diagnostics else None
    a.nVars, a.vars = len(variables), table
    L = lib()
    L.orc_upwind_run.restype = C.c_int
    err = L.orc_upwind_run(C.byref(a))
    if err:
        raise RuntimeError("orc_upwind_run: bad argument (%d)" % err)
    return out
