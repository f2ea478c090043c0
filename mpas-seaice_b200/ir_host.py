"""ctypes mirror of include/ir_b200.h: the incremental-remapping transport on the device.

Host-side counterpart of seaice_run_advection_incremental_remap
(src/shared/mpas_seaice_advection_incremental_remap.F:2338) for non-Fortran hosts: the arrays are numpy arrays in the
layout c_loc() of the MPAS pool arrays has (C order with the dimensions reversed, the extra slot n+1).  There is no
CPU fallback: without the CUDA library (or a device) construction fails.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IR_B200_LIB", os.path.join(_HERE, "csrc", "libir_b200.so"))

EXPORTS = ("ir_init_geometry", "ir_create", "ir_set_tracers", "ir_run", "ir_set_checks", "ir_fetch_check_report",
           "ir_fetch_conservation_sums", "ir_fetch_diagnostics", "ir_fetch_tracer_field", "ir_normal_vectors",
           "ir_set_upwind_mesh", "ir_run_upwind", "ir_fetch_upwind_fluxes", "ir_release_host_memory",
           "ir_last_run_ms", "ir_last_kernel_ms", "ir_launch_count", "ir_destroy", "ir_last_error_string")
GEOM_NAMES = ("x", "y", "xx", "xy", "yy", "xxx", "xxy", "xyy", "yyy", "xxxx", "xxxy", "xxyy", "xyyy", "yyyy")

IR_OK, IR_ERR_ARGUMENT, IR_ERR_CUDA, IR_ERR_STATE, IR_ERR_MESH = 0, 1, 2, 3, 4
IR_ERR_NEGATIVE_MASS_QP, IR_ERR_NEGATIVE_MASS, IR_ERR_PARALLEL_EDGES, IR_ERR_TOO_MANY_TRIANGLES = 10, 11, 12, 13
IR_ERR_CONSERVATION, IR_ERR_MONOTONICITY = 14, 15


class IrError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ir_b200 error %d: %s" % (code, message))
        self.code = code


class ir_mesh_desc(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nEdges", "maxEdges", "vertexDegree",
                                        "nCategories", "nQuadPoints", "on_a_sphere", "rotate_cartesian_grid")]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge",
                                             "verticesOnEdge", "areaCell", "dcEdge", "coeffs_reconstruct",
                                             "transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "xVertexOnEdge",
                                             "yVertexOnEdge", "remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap")]
                + [("geomAvgCell", C.c_void_p * 14)])


class ir_geometry_in(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nCellsSolve", "nVertices", "nEdges", "maxEdges", "vertexDegree",
                                        "on_a_sphere", "rotate_cartesian_grid")]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "verticesOnCell", "cellsOnEdge", "verticesOnEdge",
                                             "edgesOnVertex", "xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex",
                                             "xEdge", "yEdge", "zEdge", "dcEdge", "dvEdge")])


class ir_geometry_out(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "remapEdge",
                                           "cellsOnEdgeRemap", "edgesOnEdgeRemap", "xVertexOnEdge", "yVertexOnEdge",
                                           "minLengthEdgesOnVertex")]
                + [("geomAvgCell", C.c_void_p * 14)])


class ir_tracer_desc(C.Structure):
    _fields_ = [("nLayers", C.c_int), ("parent", C.c_int), ("volumeLike", C.c_int), ("array", C.c_void_p)]


class ir_normals_in(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nCells", "nVertices", "nVerticesSolve", "nEdges", "maxEdges", "vertexDegree",
                                        "on_a_sphere", "rotate_cartesian_grid", "remove_metric_terms")]
                + [("sphere_radius", C.c_double)]
                + [(n, C.c_void_p) for n in ("nEdgesOnCell", "edgesOnCell", "verticesOnEdge", "cellsOnEdge", "edgesOnVertex",
                                             "interiorVertex", "xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex",
                                             "xEdge", "yEdge", "zEdge")])


class ir_normals_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("normalVectorPolygon", "normalVectorTriangle", "latCellRotated", "latVertexRotated")]


class ir_upwind_var(C.Structure):
    _fields_ = [("parent", C.c_int), ("volumeLike", C.c_int), ("childMinimum", C.c_double), ("array", C.c_void_p)]


class ir_check_report(C.Structure):
    _fields_ = [("conservationViolated", C.c_int), ("consTracer", C.c_int), ("consCategory", C.c_int), ("consLayer", C.c_int),
                ("sumInit", C.c_double), ("sumFinal", C.c_double),
                ("monotonicityViolated", C.c_int), ("monoTracer", C.c_int), ("monoCategory", C.c_int), ("monoLayer", C.c_int),
                ("monoCell", C.c_int), ("newValue", C.c_double), ("bound", C.c_double), ("tolerance", C.c_double)]


_libs = {}


def load(path=None):
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise IrError(IR_ERR_STATE, "%s is missing: build it with `python __graft_entry__.py` (no CPU fallback)" % path)
        L = C.CDLL(path)
        L.ir_last_error_string.restype = C.c_char_p
        for name in EXPORTS[:-1]:
            getattr(L, name).restype = C.c_int
        _libs[path] = L
    return _libs[path]


def _ptr(a, dtype):
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], (getattr(a, "dtype", None), dtype)
    return a.ctypes.data


def init_geometry(mesh, irf, n_cells_solve=None, rotate=False, device=-1, lib_path=None):
    """The incremental_remap pool arrays (what seaice_init_advection_incremental_remap computes,
    incremental_remap.F:446-711) from the mesh-file arrays, on the device.  Returns the dict IrTransport takes."""
    L = load(lib_path)
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    gi, go = ir_geometry_in(), ir_geometry_out()
    gi.nCells, gi.nVertices, gi.nEdges, gi.maxEdges, gi.vertexDegree = nC, nV, nE, M, mesh.vertexDegree
    gi.nCellsSolve = nC if n_cells_solve is None else int(n_cells_solve)
    gi.on_a_sphere, gi.rotate_cartesian_grid = int(bool(mesh.on_a_sphere)), int(bool(rotate))
    for name in ("nEdgesOnCell", "edgesOnCell", "verticesOnCell", "cellsOnEdge"):
        setattr(gi, name, _ptr(mesh[name], np.int32))
    for name in ("verticesOnEdge", "edgesOnVertex"):
        setattr(gi, name, _ptr(irf[name], np.int32))
    for name in ("xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex", "dcEdge", "dvEdge"):
        setattr(gi, name, _ptr(mesh[name], np.float64))
    for name in ("xEdge", "yEdge", "zEdge"):
        setattr(gi, name, _ptr(irf[name], np.float64))
    out = dict(transGlobalToCell=np.zeros((max(nC, 1), 3, 3)),
               xVertexOnCell=np.zeros((nC + 1, M)), yVertexOnCell=np.zeros((nC + 1, M)),
               remapEdge=np.zeros(nE + 1, np.int32),
               cellsOnEdgeRemap=np.zeros((nE + 1, 6), np.int32), edgesOnEdgeRemap=np.zeros((nE + 1, 6), np.int32),
               xVertexOnEdge=np.zeros((nE + 1, 8)), yVertexOnEdge=np.zeros((nE + 1, 8)),
               minLengthEdgesOnVertex=np.zeros(nV + 1))
    for name, arr in out.items():
        setattr(go, name, arr.ctypes.data)
    geom = {n: np.zeros(nC + 1) for n in GEOM_NAMES}
    for k, n in enumerate(GEOM_NAMES):
        go.geomAvgCell[k] = geom[n].ctypes.data
    rc = L.ir_init_geometry(C.byref(gi), C.byref(go), C.c_int(device))
    if rc != IR_OK:
        raise IrError(rc, L.ir_last_error_string().decode())
    out["geomAvg"] = geom
    return out


def normal_vectors(mesh, edges, interior_vertex, rotate=True, remove_metric_terms=True, triangles=True,
                   n_vertices_solve=None, device=-1, lib_path=None):
    """seaice_normal_vectors (src/shared/mpas_seaice_mesh.F:703) on the device: ir_normal_vectors.  ``edges`` holds
    verticesOnEdge, edgesOnVertex, xEdge, yEdge, zEdge (mesh-file arrays; irmesh.ir_fields for generated meshes).
    Returns dict(normalVectorPolygon (nCells+1, maxEdges, 2), latCellRotated[, normalVectorTriangle (nVertices+1,
    vertexDegree, 2), latVertexRotated]); ``triangles=False`` is seaice_normal_vectors_polygon alone."""
    L = load(lib_path)
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    a = ir_normals_in()
    a.nCells, a.nVertices, a.nEdges, a.maxEdges, a.vertexDegree = nC, nV, nE, M, D
    a.nVerticesSolve = nV if n_vertices_solve is None else int(n_vertices_solve)
    a.on_a_sphere = int(bool(mesh.on_a_sphere))
    a.rotate_cartesian_grid, a.remove_metric_terms = int(bool(rotate)), int(bool(remove_metric_terms))
    a.sphere_radius = float(getattr(mesh, "sphere_radius", 0.0) or 0.0)
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
    o = ir_normals_out()
    o.normalVectorPolygon, o.latCellRotated = out["normalVectorPolygon"].ctypes.data, out["latCellRotated"].ctypes.data
    if triangles:
        out.update(normalVectorTriangle=np.zeros((nV + 1, D, 2)), latVertexRotated=np.zeros(nV + 1))
        o.normalVectorTriangle, o.latVertexRotated = out["normalVectorTriangle"].ctypes.data, out["latVertexRotated"].ctypes.data
    rc = L.ir_normal_vectors(C.byref(a), C.byref(o), C.c_int(device))
    if rc != IR_OK:
        raise IrError(rc, L.ir_last_error_string().decode())
    return out


def interior_edge(mesh, n_edges_solve=None):
    """interiorEdge of the boundary pool (interior_edges, src/shared/mpas_seaice_mesh.F:567): 1 for the first nEdgesSolve
    edges that have a cell of the block on both sides."""
    nC, nE = mesh.nCells, mesh.nEdges
    nES = nE if n_edges_solve is None else int(n_edges_solve)
    out = np.zeros(nE + 1, np.int32)
    coe = mesh.cellsOnEdge[:nES]
    out[:nES] = ((coe[:, 0] <= nC) & (coe[:, 1] <= nC)).astype(np.int32)
    return out


class IrTransport:
    """One block's transport.  ``mesh``: meshgen.Mesh (or any mapping with the MPAS mesh-pool names); ``irf``: the
    mesh-file / framework arrays of irmesh.ir_fields; ``geom``: the incremental_remap pool arrays
    (xVertexOnCell ... geomAvg) as seaice_init_advection_incremental_remap leaves them."""

    def __init__(self, mesh, irf, geom, n_categories, n_quad_points=6, n_cells_solve=None, rotate=False, device=-1,
                 lib_path=None):
        self._L = load(lib_path)
        self._h = C.c_void_p()
        self.mesh = mesh
        nC = mesh.nCells
        m = ir_mesh_desc()
        m.nCells, m.nVertices, m.nEdges, m.maxEdges, m.vertexDegree = nC, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
        m.nCellsSolve = nC if n_cells_solve is None else int(n_cells_solve)
        m.nCategories, m.nQuadPoints = int(n_categories), int(n_quad_points)
        m.on_a_sphere, m.rotate_cartesian_grid = int(bool(mesh.on_a_sphere)), int(bool(rotate))
        self._keep = [mesh, irf, geom]
        for name in ("nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "cellsOnEdge"):
            setattr(m, name, _ptr(mesh[name], np.int32))
        m.verticesOnEdge = _ptr(irf["verticesOnEdge"], np.int32)
        m.areaCell, m.dcEdge = _ptr(mesh["areaCell"], np.float64), _ptr(mesh["dcEdge"], np.float64)
        m.coeffs_reconstruct = _ptr(irf["coeffs_reconstruct"], np.float64)
        m.transGlobalToCell = _ptr(geom["transGlobalToCell"], np.float64) if mesh.on_a_sphere else None
        for name in ("xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge"):
            setattr(m, name, _ptr(geom[name], np.float64))
        for name in ("remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap"):
            setattr(m, name, _ptr(geom[name], np.int32))
        for k, n in enumerate(GEOM_NAMES):
            m.geomAvgCell[k] = _ptr(geom["geomAvg"][n], np.float64)
        self.n_categories, self.n_quad_points = int(n_categories), int(n_quad_points)
        self._check(self._L.ir_create(C.byref(self._h), C.byref(m), C.c_int(device)))
        self._table = None

    def _check(self, rc):
        if rc != IR_OK:
            raise IrError(rc, self._L.ir_last_error_string().decode())

    def _make_table(self, tracers):
        """tracers: sequence of objects with .array (nCells+1, nCategories, nLayers), .parent (index or None),
        .volume_like"""
        table = (ir_tracer_desc * len(tracers))()
        for i, t in enumerate(tracers):
            assert t.array.shape[:2] == (self.mesh.nCells + 1, self.n_categories)
            table[i].nLayers = t.array.shape[2]
            table[i].parent = -1 if t.parent is None else int(t.parent)
            table[i].volumeLike = int(bool(t.volume_like))
            table[i].array = _ptr(t.array, np.float64)
        return table

    def set_tracers(self, tracers):
        self._table = self._make_table(tracers)
        self._check(self._L.ir_set_tracers(self._h, C.c_int(len(tracers)), self._table))

    def run(self, tracers, u, v, dt, check=True):
        """One step IN PLACE on the tracer arrays; returns the ir_run code (raises on an abort condition if check)."""
        nV = self.mesh.nVertices
        assert u.shape == (nV + 1,) and v.shape == (nV + 1,)
        table = self._make_table(tracers)
        rc = self._L.ir_run(self._h, C.c_int(len(tracers)), table, C.c_void_p(_ptr(u, np.float64)),
                            C.c_void_p(_ptr(v, np.float64)), C.c_double(dt))
        if check or rc in (IR_ERR_ARGUMENT, IR_ERR_CUDA, IR_ERR_STATE, IR_ERR_MESH):
            self._check(rc)
        return rc

    def set_checks(self, conservation=0, monotonicity=0):
        """config_conservation_check (1: sums and check, 2: sums only) / config_monotonicity_check of the next runs."""
        self._check(self._L.ir_set_checks(self._h, C.c_int(int(conservation)), C.c_int(int(monotonicity))))

    def check_report(self):
        """ir_check_report of the last run as a dict."""
        rep = ir_check_report()
        self._check(self._L.ir_fetch_check_report(self._h, C.byref(rep)))
        return {name: getattr(rep, name) for name, _ in ir_check_report._fields_}

    def conservation_sums(self, tracer_index, n_layers):
        """(sumInit, sumFinal) of one tracer after a run with the conservation check on: (nCategories, nLayers) each."""
        a, b = np.zeros((self.n_categories, n_layers)), np.zeros((self.n_categories, n_layers))
        self._check(self._L.ir_fetch_conservation_sums(self._h, C.c_int(tracer_index), C.c_void_p(a.ctypes.data),
                                                       C.c_void_p(b.ctypes.data)))
        return a, b

    # ---- config_advection_type = 'upwind' (seaice_run_advection_upwind, advection_upwind.F:385)
    def set_upwind_mesh(self, interior_edge_mask, dv_edge, normal_vector_edge):
        nC, nE, M = self.mesh.nCells, self.mesh.nEdges, self.mesh.maxEdges
        assert interior_edge_mask.shape == (nE + 1,) and dv_edge.shape == (nE + 1,) and normal_vector_edge.shape == (nC + 1, M, 2)
        self._check(self._L.ir_set_upwind_mesh(self._h, C.c_void_p(_ptr(interior_edge_mask, np.int32)),
                                               C.c_void_p(_ptr(dv_edge, np.float64)),
                                               C.c_void_p(_ptr(normal_vector_edge, np.float64))))

    def run_upwind(self, variables, u, v, dt):
        """One upwind step IN PLACE; ``variables``: objects with .array (nCells+1, nCategories), .parent (index or
        None), .volume_like, .child_minimum -- the rows of the reference's tracerConnectivities table, in order."""
        nV = self.mesh.nVertices
        assert u.shape == (nV + 1,) and v.shape == (nV + 1,)
        table = (ir_upwind_var * len(variables))()
        for i, var in enumerate(variables):
            assert var.array.shape == (self.mesh.nCells + 1, self.n_categories)
            table[i].parent = -1 if var.parent is None else int(var.parent)
            table[i].volumeLike = int(bool(var.volume_like))
            table[i].childMinimum = float(var.child_minimum)
            table[i].array = _ptr(var.array, np.float64)
        self._check(self._L.ir_run_upwind(self._h, C.c_int(len(variables)), table, C.c_void_p(_ptr(u, np.float64)),
                                          C.c_void_p(_ptr(v, np.float64)), C.c_double(dt)))

    def upwind_fluxes(self, var_index):
        """(<variable>EdgeFlux (nEdges+1, nCategories), edgeVelocity (nEdges+1)) of the last upwind step."""
        nE = self.mesh.nEdges
        flux, vel = np.zeros((nE + 1, self.n_categories)), np.zeros(nE + 1)
        self._check(self._L.ir_fetch_upwind_fluxes(self._h, C.c_int(var_index), C.c_void_p(flux.ctypes.data),
                                                   C.c_void_p(vel.ctypes.data)))
        return flux, vel

    def diagnostics(self, n_mass_layers=1):
        nE, nQ, nK = self.mesh.nEdges, self.n_quad_points, self.n_categories
        out = dict(xTriangle=np.zeros((nE, 6, nQ)), yTriangle=np.zeros((nE, 6, nQ)), triangleArea=np.zeros((nE, 6)),
                   iCellTriangle=np.zeros((nE, 6), np.int32), maskEdge=np.zeros(nE, np.int32),
                   edgeFluxMass=np.zeros((nE, nK, n_mass_layers)))
        self._check(self._L.ir_fetch_diagnostics(self._h, *[C.c_void_p(out[k].ctypes.data) for k in
                                                            ("xTriangle", "yTriangle", "triangleArea", "iCellTriangle",
                                                             "maskEdge", "edgeFluxMass")]))
        return out

    FIELDS = dict(center=0, xGrad=1, yGrad=2, xBarycenter=3, yBarycenter=4, massTracerProduct=5, edgeFlux=6)

    def tracer_field(self, which, tracer_index, n_layers):
        """A work field of the last step (see ir_fetch_tracer_field): (nCells+1 | nEdges+1, nCategories, nLayers)."""
        n = (self.mesh.nEdges if which == "edgeFlux" else self.mesh.nCells) + 1
        out = np.zeros((n, self.n_categories, n_layers))
        self._check(self._L.ir_fetch_tracer_field(self._h, C.c_int(self.FIELDS[which]), C.c_int(tracer_index),
                                                  C.c_void_p(out.ctypes.data)))
        return out

    def release_host_memory(self):
        """Undo the page-locking of host arrays done under IR_B200_PIN_HOST (before they are freed)."""
        self._check(self._L.ir_release_host_memory(self._h))

    def last_run_ms(self):
        ms = C.c_float(0)
        self._check(self._L.ir_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        """Device time of the last run by kernel: dict(prepare, reconstruct, triangles, fluxes, update) in ms."""
        ms = (C.c_float * 5)()
        self._check(self._L.ir_last_kernel_ms(self._h, ms))
        return dict(zip(("prepare", "reconstruct", "triangles", "fluxes", "update"), (float(x) for x in ms)))

    def launch_count(self):
        n = C.c_longlong(0)
        self._check(self._L.ir_launch_count(self._h, C.byref(n)))
        return n.value

    def destroy(self):
        if self._h:
            self._L.ir_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class TracerHalo:
    """Owner -> halo update of the tracer arrays between the ranks of a decomposed run: what
    seaice_update_tracer_halo does after the transport (incremental_remap.F:2705-2712).  Host-side, over the
    process group ``dist`` (torch.distributed with a backend that moves CPU tensors, e.g. gloo), one message per
    neighbour carrying every tracer.  ``lists``: partition.cell_exchange_lists of this rank's block."""

    def __init__(self, lists, dist):
        self.nbrs, self.send_off, self.send_idx, self.recv_off, self.recv_idx = lists
        self.dist = dist

    def update(self, tracers):
        import torch
        dist = self.dist
        ops, recv_bufs = [], []
        for i, q in enumerate(self.nbrs):
            s = self.send_idx[self.send_off[i]:self.send_off[i + 1]].astype(np.int64) - 1
            r = self.recv_idx[self.recv_off[i]:self.recv_off[i + 1]].astype(np.int64) - 1
            if s.size:
                out = torch.from_numpy(np.concatenate([t.array[s].reshape(-1) for t in tracers]))
                ops.append(dist.P2POp(dist.isend, out, int(q)))
            if r.size:
                n = sum(r.size * t.array.shape[1] * t.array.shape[2] for t in tracers)
                buf = torch.empty(n, dtype=torch.float64)
                ops.append(dist.P2POp(dist.irecv, buf, int(q)))
                recv_bufs.append((r, buf))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for r, buf in recv_bufs:
            a, o = buf.numpy(), 0
            for t in tracers:
                n = r.size * t.array.shape[1] * t.array.shape[2]
                t.array[r] = a[o:o + n].reshape((r.size,) + t.array.shape[1:])
                o += n
