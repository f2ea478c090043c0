"""Edge connectivity and normal vectors for the weak operators, for synthetic hosts.

In MPAS-Seaice ``verticesOnEdge`` / ``edgesOnVertex`` come from the mesh file and the normal vectors
from seaice_normal_vectors (reference: src/shared/mpas_seaice_mesh.F:703-2007, weak schemes only --
out of scope of this library).  The generators of meshgen.py do not emit
them, so this module derives the connectivity; on planar meshes the normals follow the reference's
normal_vectors_planar_polygon / _triangle (mesh.F:858-1024), on the sphere they are geometrically
sensible unit normals in the local tangent planes (the reference's great-circle construction,
mesh.F:1038-1744, is not restated).  Operator parity (device vs oracle) does not depend on how the normals were
made: both sides read the same arrays.
"""
from __future__ import annotations

import numpy as np

from . import variational_init


def edge_connectivity(mesh):
    """verticesOnEdge (nEdges+1, 2) and edgesOnVertex (nVertices+1, vertexDegree), 1-based, junk = n+1.
    Edge k of a cell joins two consecutive vertices of the cell; which two is decided by the cell on the
    other side (an edge's vertices are the vertices its two cells share)."""
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    voc, eoc, n_on = mesh.verticesOnCell, mesh.edgesOnCell, mesh.nEdgesOnCell
    voe = np.full((nE + 1, 2), nV + 1, dtype=np.int32)
    # meshgen convention: edge k of a cell joins verticesOnCell(k) and verticesOnCell(k+1) (cyclic); every
    # cell that lists an edge must agree on its two vertices
    es, los, his = [], [], []
    for k in range(M):
        c = np.nonzero(n_on[:nC] > k)[0]
        kk = (k + 1) % n_on[c]
        a = voc[c, k].astype(np.int64)
        b = voc[c, kk].astype(np.int64)
        es.append(eoc[c, k].astype(np.int64) - 1)
        los.append(np.minimum(a, b))
        his.append(np.maximum(a, b))
    es, los, his = np.concatenate(es), np.concatenate(los), np.concatenate(his)
    uniq, first = np.unique(es, return_index=True)
    lo_e = np.full(nE, -1, dtype=np.int64)
    hi_e = np.full(nE, -1, dtype=np.int64)
    lo_e[uniq], hi_e[uniq] = los[first], his[first]
    assert np.all(lo_e[es] == los) and np.all(hi_e[es] == his), "edgesOnCell is not aligned with verticesOnCell"
    has = lo_e >= 0
    voe[:nE][has, 0] = lo_e[has].astype(np.int32)
    voe[:nE][has, 1] = hi_e[has].astype(np.int32)
    # edgesOnVertex: slot s = the edge shared by cellsOnVertex(s) and cellsOnVertex(s+1) (where both exist)
    cov = mesh.cellsOnVertex
    coe = mesh.cellsOnEdge
    eov = np.full((nV + 1, D), nE + 1, dtype=np.int32)
    c1 = np.minimum(coe[:nE, 0], coe[:nE, 1]).astype(np.int64)
    c2 = np.maximum(coe[:nE, 0], coe[:nE, 1]).astype(np.int64)
    order = np.argsort(c1 * (nC + 2) + c2, kind="stable")
    sorted_key = (c1 * (nC + 2) + c2)[order]
    for s in range(D):
        a = cov[:nV, s].astype(np.int64)
        b = cov[:nV, (s + 1) % D].astype(np.int64)
        k = np.minimum(a, b) * (nC + 2) + np.maximum(a, b)
        pos = np.searchsorted(sorted_key, k)
        pos = np.minimum(pos, nE - 1)
        hit = (sorted_key[pos] == k) & (a <= nC) & (b <= nC) & (a >= 1) & (b >= 1)
        eov[:nV, s][hit] = (order[pos][hit] + 1).astype(np.int32)
    return voe, eov


def _planar_fields(mesh, voe, eov):
    """Planar meshes: the reference's own definitions, normal_vectors_planar_polygon / _triangle
    (src/shared/mpas_seaice_mesh.F:858-943, 957-1024), with xEdge / yEdge = the edge midpoint."""
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    xv, yv = mesh.xVertex, mesh.yVertex
    xe = np.zeros(nE + 1)
    ye = np.zeros(nE + 1)
    v1, v2 = voe[:nE, 0] - 1, voe[:nE, 1] - 1
    xe[:nE] = 0.5 * (xv[v1] + xv[v2])
    ye[:nE] = 0.5 * (yv[v1] + yv[v2])
    nvp = np.zeros((nC + 1, M, 2))
    for k in range(M):
        c = np.nonzero(mesh.nEdgesOnCell[:nC] > k)[0]
        e = mesh.edgesOnCell[c, k] - 1
        a, b = voe[e, 0] - 1, voe[e, 1] - 1
        tx, ty = xv[b] - xv[a], yv[b] - yv[a]
        tmag = np.sqrt(tx ** 2 + ty ** 2)
        tx, ty = tx / tmag, ty / tmag
        nx, ny = xe[e] - mesh.xCell[c], ye[e] - mesh.yCell[c]
        flip = (nx * ty - ny * tx) < 0.0
        tx = np.where(flip, -tx, tx)
        ty = np.where(flip, -ty, ty)
        nvp[c, k, 0] = ty
        nvp[c, k, 1] = -tx
    nvt = np.zeros((nV + 1, D, 2))
    interior = variational_init.interior_vertex(mesh)[:nV] == 1
    v = np.nonzero(interior)[0]
    for s in range(D):
        e = eov[v, s] - 1
        dx, dy = xe[e] - xv[v], ye[e] - yv[v]
        nvt[v, s, 0] = dx / np.sqrt(dx ** 2 + dy ** 2)
        nvt[v, s, 1] = dy / np.sqrt(dx ** 2 + dy ** 2)
    return dict(verticesOnEdge=voe, edgesOnVertex=eov, normalVectorPolygon=nvp, normalVectorTriangle=nvt,
                latCellRotated=np.zeros(nC + 1), latVertexRotated=np.zeros(nV + 1))


def weak_fields(mesh):
    """dict(verticesOnEdge, edgesOnVertex, normalVectorPolygon (nCells+1, maxEdges, 2), normalVectorTriangle
    (nVertices+1, vertexDegree, 2), latCellRotated, latVertexRotated) in Registry layouts."""
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    voe, eov = edge_connectivity(mesh)
    on_sphere = bool(mesh.on_a_sphere)
    if not on_sphere:
        return _planar_fields(mesh, voe, eov)
    xl, yl = variational_init.local_coords(mesh, rotate=on_sphere)       # vertices in the cell's tangent plane
    n_on = mesh.nEdgesOnCell
    nvp = np.zeros((nC + 1, M, 2))
    # which neighbour vertex closes edge k: decided exactly as in edge_connectivity (re-derive from verticesOnEdge)
    voc = mesh.verticesOnCell
    for k in range(M):
        valid = n_on[:nC] > k
        c = np.nonzero(valid)[0]
        e = mesh.edgesOnCell[c, k]
        other = np.where(voe[e - 1, 0] == voc[c, k], voe[e - 1, 1], voe[e - 1, 0])
        # slot of the other vertex inside the cell
        slot = np.argmax(voc[c, :] == other[:, None], axis=1)
        dx = xl[c, slot] - xl[c, k]
        dy = yl[c, slot] - yl[c, k]
        nx, ny = dy, -dx
        # outward: pointing away from the cell centre (the origin of the local coordinates)
        mx, my = 0.5 * (xl[c, slot] + xl[c, k]), 0.5 * (yl[c, slot] + yl[c, k])
        flip = (nx * mx + ny * my) < 0.0
        nx = np.where(flip, -nx, nx)
        ny = np.where(flip, -ny, ny)
        ln = np.sqrt(nx * nx + ny * ny)
        nvp[c, k, 0] = nx / ln
        nvp[c, k, 1] = ny / ln
    # dual triangle: edge s of vertex v is crossed by the dual edge joining the two cells of that edge
    if on_sphere:
        px, py, pz = -mesh.zCell, mesh.yCell, mesh.xCell         # rotated frame, as local_coords uses
        vx, vy, vz = -mesh.zVertex, mesh.yVertex, mesh.xVertex
        r = np.sqrt(vx * vx + vy * vy + vz * vz)
        r[r == 0] = 1.0
        ux, uy, uz = vx / r, vy / r, vz / r
        ex, ey, ez = -uy, ux, np.zeros_like(ux)                   # east
        en = np.sqrt(ex * ex + ey * ey)
        en[en == 0] = 1.0
        ex, ey = ex / en, ey / en
        nx3, ny3, nz3 = uy * ez - uz * ey, uz * ex - ux * ez, ux * ey - uy * ex   # north = up x east

        def local(v, c):
            dx, dy, dz = px[c] - vx[v], py[c] - vy[v], pz[c] - vz[v]
            return dx * ex[v] + dy * ey[v] + dz * ez[v], dx * nx3[v] + dy * ny3[v] + dz * nz3[v]
    else:
        def local(v, c):
            return mesh.xCell[c] - mesh.xVertex[v], mesh.yCell[c] - mesh.yVertex[v]
    nvt = np.zeros((nV + 1, D, 2))
    v = np.arange(nV)
    for s in range(D):
        e = eov[:nV, s]
        ok = e <= mesh.nEdges
        vv = v[ok]
        ca = mesh.cellsOnEdge[e[ok] - 1, 0] - 1
        cb = mesh.cellsOnEdge[e[ok] - 1, 1] - 1
        good = (ca < nC) & (cb < nC) & (ca >= 0) & (cb >= 0)
        vv, ca, cb = vv[good], ca[good], cb[good]
        ax, ay = local(vv, ca)
        bx, by = local(vv, cb)
        dx, dy = bx - ax, by - ay
        nx, ny = dy, -dx
        mx, my = 0.5 * (ax + bx), 0.5 * (ay + by)
        flip = (nx * mx + ny * my) < 0.0
        nx = np.where(flip, -nx, nx)
        ny = np.where(flip, -ny, ny)
        ln = np.sqrt(nx * nx + ny * ny)
        nvt[vv, s, 0] = nx / ln
        nvt[vv, s, 1] = ny / ln
    lat_c = np.zeros(nC + 1)
    lat_v = np.zeros(nV + 1)
    if on_sphere:
        R = mesh.sphere_radius
        lat_c[:nC] = np.arcsin(np.clip(mesh.xCell[:nC] / R, -1.0, 1.0))      # rotated pole: z' = x
        lat_v[:nV] = np.arcsin(np.clip(mesh.xVertex[:nV] / R, -1.0, 1.0))
    return dict(verticesOnEdge=voe, edgesOnVertex=eov, normalVectorPolygon=nvp, normalVectorTriangle=nvt,
                latCellRotated=lat_c, latVertexRotated=lat_v)
