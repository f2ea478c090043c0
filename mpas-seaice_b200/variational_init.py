"""Host-side (numpy, vectorised) mirror of the cheap parts of seaice_init_velocity_solver_variational
(reference: src/shared/mpas_seaice_velocity_solver_variational.F:108-344): metric terms,
cellVerticesAtVertex, local tangent-plane coordinates, interior vertices and the 'original'
denominator.  The expensive part -- the Wachspress basis -- is computed on the device by
evp_precompute_wachspress() from the local coordinates produced here.

In a real MPAS-Seaice run these arrays come from the Fortran init and this module is not needed; it
exists so that bench.py and the synthetic hosts can set up 10 M-cell problems in seconds.  The
arithmetic order follows the cited lines (checked bit-for-bit against the oracle in the tests).
"""
from __future__ import annotations

import numpy as np


def metric_terms(mesh, rotate=True, include=True):
    """seaice_calc_variational_metric_terms (variational_shared.F:293-358)"""
    nV = mesh.nVertices
    out = np.zeros(nV + 1)
    if include:
        z = mesh.xVertex[:nV] if rotate else mesh.zVertex[:nV]      # rotation (x,y,z) -> (-z, y, x): zp = x
        # NOT np.tan(np.arcsin(.)): numpy's SIMD transcendental loops give position-dependent last bits,
        # and a block-local evaluation must reproduce the global one bit for bit
        from . import host
        out[:nV] = host.host_metric_terms(z, mesh.sphere_radius)
    return out


def cell_vertices_at_vertex(mesh):
    """seaice_cell_vertices_at_vertex (src/shared/mpas_seaice_mesh.F:632-685); the reference keeps the
    LAST matching slot, 0 if the cell does not hold the vertex."""
    nV, nC, D, M = mesh.nVertices, mesh.nCells, mesh.vertexDegree, mesh.maxEdges
    cov = mesh.cellsOnVertex
    out = np.zeros((nV + 1, D), dtype=np.int32)
    vid = np.arange(1, nV + 1, dtype=np.int32)
    for k in range(D):
        c = cov[:nV, k] - 1                       # junk cell (nC) has nEdgesOnCell = 0
        n = mesh.nEdgesOnCell[c]
        for j in range(M):
            hit = (j < n) & (mesh.verticesOnCell[c, j] == vid)
            out[:nV, k][hit] = j + 1
    return out


def interior_vertex(mesh):
    """interior_vertices (mesh.F:423-488)"""
    out = np.zeros(mesh.nVertices + 1, dtype=np.int32)
    cov = mesh.cellsOnVertex[:mesh.nVertices]
    out[:mesh.nVertices] = np.all((cov >= 1) & (cov <= mesh.nCells), axis=1)
    return out


def boundary_source_local(index_to_id, boundary_type, boundary_source):
    """init_special_boundaries_velocity / init_special_boundaries_tracers (special_boundaries.F:83-150, 164-250):
    ``vertexBoundarySourceLocal`` (what evp_create's special-boundary arrays take, include/evp_b200.h) /
    ``tracerBoundarySourceLocal`` from the global IDs in the stream arrays ``vertexBoundarySource`` /
    ``tracerBoundarySource``.  Arrays carry the extra slot; entities of type 0 get 0.  Like the reference's table
    (allocated with nVertices / nCells slots) this serves a block whose IDs are a permutation of 1..n; anything else
    raises instead of indexing out of bounds."""
    ids = np.asarray(index_to_id)
    n = len(ids) - 1
    btype, src = np.asarray(boundary_type), np.asarray(boundary_source)
    assert btype.shape == (n + 1,) and src.shape == (n + 1,)
    special = np.flatnonzero(btype[:n] != 0)
    if np.any(ids[:n] < 1) or np.any(ids[:n] > n) or np.any(src[special] < 1) or np.any(src[special] > n):
        raise ValueError("global IDs outside 1..n: the reference's globalToLocalID table would be indexed out of bounds")
    g2l = np.zeros(n + 1, dtype=np.int32)
    g2l[ids[:n]] = np.arange(1, n + 1, dtype=np.int32)
    out = np.zeros(n + 1, dtype=np.int32)
    out[special] = g2l[src[special]]
    return out


def set_special_boundaries_tracers(boundary_type, source_local, *category_arrays):
    """seaice_set_special_boundaries_tracers (special_boundaries.F:415-485), IN PLACE on the category tracers
    (first dimension nCells+1): type 1 cells are emptied, type 2 cells take their source cell's values -- in cell order,
    so a source the loop changed earlier is read changed (the reference's in-place semantics)."""
    btype = np.asarray(boundary_type)
    for i in np.flatnonzero(btype[:-1] != 0):
        for a in category_arrays:
            if btype[i] == 1:
                a[i] = 0.0
            elif btype[i] == 2:
                a[i] = a[int(source_local[i]) - 1]


def land_ice_mask_vertex(mesh, land_ice_mask, n_vertices_solve=None):
    """init_ice_shelve_vertex_mask (velocity_solver.F:481-544): 1 at the vertices of the owned range that touch a cell
    under an ice shelf (landIceMask == 1; the junk slot nCells+1 of the mask counts like any cell), 0 elsewhere --
    the ``landIceMaskVertex`` of evp_mesh_ext (include/evp_b200.h)."""
    nV = mesh.nVertices
    nVs = nV if n_vertices_solve is None else int(n_vertices_solve)
    land = np.asarray(land_ice_mask)
    assert land.shape == (mesh.nCells + 1,)
    out = np.zeros(nV + 1, dtype=np.int32)
    out[:nVs] = np.any(land[mesh.cellsOnVertex[:nVs] - 1] == 1, axis=1)
    return out


def dynamically_locked_cells_mask(mesh, interior_vertex_mask):
    """dynamically_locked_cell_mask (velocity_solver.F:402-467): 1 at the cells none of whose vertices is an interior
    vertex (their ice cannot move), the mask the regional statistics read."""
    nC = mesh.nCells
    out = np.zeros(nC + 1, dtype=np.int32)
    valid = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
    voc = np.where(valid, mesh.verticesOnCell[:nC] - 1, mesh.nVertices)
    out[:nC] = ~np.any(valid & (np.asarray(interior_vertex_mask)[voc] == 1), axis=1)
    return out


def local_coords(mesh, rotate=True):
    """seaice_calc_local_coords (variational_shared.F:42-279) with
    seaice_project_3D_vector_onto_local_2D (mesh.F:2021-2061, 2272-2332)."""
    nC, M = mesh.nCells, mesh.maxEdges
    xl = np.zeros((nC + 1, M))
    yl = np.zeros((nC + 1, M))
    n = mesh.nEdgesOnCell[:nC]
    voc = mesh.verticesOnCell[:nC] - 1
    if not mesh.on_a_sphere:
        for j in range(M):
            ok = j < n
            v = voc[ok, j]
            xl[:nC][ok, j] = mesh.xVertex[v] - mesh.xCell[:nC][ok]
            yl[:nC][ok, j] = mesh.yVertex[v] - mesh.yCell[:nC][ok]
        return xl, yl
    if rotate:
        xc, yc, zc = -mesh.zCell[:nC], mesh.yCell[:nC], mesh.xCell[:nC]
        vxr, vyr, vzr = -mesh.zVertex, mesh.yVertex, mesh.xVertex
    else:
        xc, yc, zc = mesh.xCell[:nC], mesh.yCell[:nC], mesh.zCell[:nC]
        vxr, vyr, vzr = mesh.xVertex, mesh.yVertex, mesh.zVertex
    with np.errstate(all="ignore"):
        e0, e1, e2 = -yc, xc, np.zeros(nC)
        mag = np.sqrt(e0 * e0 + e1 * e1 + e2 * e2)
        e0, e1, e2 = e0 / mag, e1 / mag, e2 / mag
        n0, n1, n2 = -xc, -yc, (xc * xc + yc * yc) / zc
        mag = np.sqrt(n0 * n0 + n1 * n1 + n2 * n2)
        n0, n1, n2 = n0 / mag, n1 / mag, n2 / mag
        south = zc < 0.0
        n0 = np.where(south, -n0, n0)
        n1 = np.where(south, -n1, n1)
        n2 = np.where(south, -n2, n2)
        eq = zc == 0.0
        n0 = np.where(eq, 0.0, n0)
        n1 = np.where(eq, 0.0, n1)
        n2 = np.where(eq, 1.0, n2)
    for j in range(M):
        ok = j < n
        v = voc[ok, j]
        xl[:nC][ok, j] = vxr[v] * e0[ok] + vyr[v] * e1[ok] + vzr[v] * e2[ok]
        yl[:nC][ok, j] = vxr[v] * n0[ok] + vyr[v] * n1[ok] + vzr[v] * n2[ok]
    return xl, yl


def init_static(mesh, rotate=None, metric=None):
    """The static ``velocity_variational`` fields except the basis arrays ('original' denominator =
    areaTriangle, variational.F:433-437) plus the local coordinates the device precompute consumes."""
    on_sphere = bool(mesh.on_a_sphere)
    rotate = on_sphere if rotate is None else rotate
    metric = on_sphere if metric is None else metric
    xl, yl = local_coords(mesh, rotate)
    den = np.zeros(mesh.nVertices + 1)
    den[:mesh.nVertices] = mesh.areaTriangle[:mesh.nVertices]
    return dict(tanLatVertexRotatedOverRadius=metric_terms(mesh, rotate, metric),
                cellVerticesAtVertex=cell_vertices_at_vertex(mesh),
                interiorVertex=interior_vertex(mesh),
                variationalDenominator=den, xLocal=xl, yLocal=yl)
