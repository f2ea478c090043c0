"""Edge connectivity and normal vectors for the weak operators, for non-Fortran hosts.

In MPAS-Seaice ``verticesOnEdge`` / ``edgesOnVertex`` come from the mesh file and the normal vectors from
seaice_normal_vectors (reference: src/shared/mpas_seaice_mesh.F:703-2007), which the Fortran host keeps
computing itself.  The generators of meshgen.py emit neither, so this module derives the connectivity (with the
MPAS orientation of verticesOnEdge against cellsOnEdge) and restates the reference's normal vectors:
planar meshes normal_vectors_planar_polygon / _triangle (mesh.F:858-1024), spherical meshes
normal_vectors_spherical_polygon_metric / _triangle_metric (mesh.F:1038-1606) with removeMetricTerms = .true.
as seaice_init_velocity_solver_weak calls them (weak.F:87-96).  tests/test_analytic_golden.py checks the result
through the weak operators against the reference's analytic operator-test fields.
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


def _sphere_fields(mesh, voe, eov):
    """Spherical meshes: normal_vectors_spherical_polygon_metric / _triangle_metric of the reference
    (src/shared/mpas_seaice_mesh.F:1038-1241, 1393-1606) with config_rotate_cartesian_grid = true and
    removeMetricTerms = .true. as seaice_init_velocity_solver_weak passes it (weak.F:87-96): every cell (vertex) is
    first rotated to lon = 0, lat = 0 of the rotated grid, where east / north are the axes of its tangent plane; the
    horizontal unit normal of the side (tangent x position) is then expressed by its eastward component and the
    signed remainder.  xEdge / yEdge / zEdge = the vertex midpoint pushed back onto the sphere.  The reference's sign
    rule relies on the MPAS orientation of verticesOnEdge against cellsOnEdge (tangent = radial x normal), which is
    established here first."""
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    R = float(mesh.sphere_radius)
    rot = lambda x, y, z: (-z, y, x)                       # seaice_grid_rotation_forward, mesh.F:2367-2379
    cx, cy, cz = rot(mesh.xCell, mesh.yCell, mesh.zCell)
    vx, vy, vz = rot(mesh.xVertex, mesh.yVertex, mesh.zVertex)
    coe = mesh.cellsOnEdge[:nE].astype(np.int64) - 1
    a, b = voe[:nE, 0].astype(np.int64) - 1, voe[:nE, 1].astype(np.int64) - 1
    ex, ey, ez = 0.5 * (vx[a] + vx[b]), 0.5 * (vy[a] + vy[b]), 0.5 * (vz[a] + vz[b])
    en = np.sqrt(ex * ex + ey * ey + ez * ez) / R
    ex, ey, ez = ex / en, ey / en, ez / en
    # orientation: (v2 - v1) . (r x (c2 - c1)) > 0
    nx_, ny_, nz_ = cx[coe[:, 1]] - cx[coe[:, 0]], cy[coe[:, 1]] - cy[coe[:, 0]], cz[coe[:, 1]] - cz[coe[:, 0]]
    tx_, ty_, tz_ = ey * nz_ - ez * ny_, ez * nx_ - ex * nz_, ex * ny_ - ey * nx_
    swap = ((vx[b] - vx[a]) * tx_ + (vy[b] - vy[a]) * ty_ + (vz[b] - vz[a]) * tz_) < 0.0
    voe = voe.copy()
    voe[:nE][swap] = voe[:nE][swap][:, ::-1]
    a, b = voe[:nE, 0].astype(np.int64) - 1, voe[:nE, 1].astype(np.int64) - 1

    def to_equator(px, py, pz, lat, lon):
        """yRotationMatrix . zRotationMatrix . p  (mesh.F:1118-1127): the point (lat, lon) goes to (R, 0, 0)."""
        cl, sl = np.cos(-lon), np.sin(-lon)
        x1, y1, z1 = cl * px - sl * py, sl * px + cl * py, pz
        ct, st = np.cos(lat), np.sin(lat)
        return ct * x1 + st * z1, y1, -st * x1 + ct * z1

    def components(gx, gy, gz, qx, qy):
        """Unit horizontal normal (gx,gy,gz) at the rotated edge position (qx,qy,.) -> (eastward component, signed
        remainder), mesh.F:1205-1225."""
        gn = np.sqrt(gx ** 2 + gy ** 2 + gz ** 2)
        gx, gy, gz = gx / gn, gy / gn, gz / gn
        east_x, east_y = -qy, qx
        east_n = np.sqrt(east_x ** 2 + east_y ** 2)
        east_x, east_y = east_x / east_n, east_y / east_n
        n1 = gx * east_x + gy * east_y + gz * 0.0
        n2 = np.copysign(1.0, gz) * np.sqrt(1.0 - np.maximum(np.minimum(n1, 1.0), -1.0) ** 2)
        return n1, n2

    lat_cell = np.arcsin(np.clip(cz[:nC] / R, -1.0, 1.0))
    lon_cell = np.arctan2(cy[:nC], cx[:nC])
    nvp = np.zeros((nC + 1, M, 2))
    for k in range(M):
        c = np.nonzero(mesh.nEdgesOnCell[:nC] > k)[0]
        e = mesh.edgesOnCell[c, k].astype(np.int64) - 1
        la, lo = lat_cell[c], lon_cell[c]
        qx, qy, qz = to_equator(ex[e], ey[e], ez[e], la, lo)
        ax_, ay_, az_ = to_equator(vx[a[e]], vy[a[e]], vz[a[e]], la, lo)
        bx_, by_, bz_ = to_equator(vx[b[e]], vy[b[e]], vz[b[e]], la, lo)
        wx, wy, wz = bx_ - ax_, by_ - ay_, bz_ - az_
        gx, gy, gz = wy * qz - wz * qy, wz * qx - wx * qz, wx * qy - wy * qx
        flip = np.where(c == coe[e, 1], -1.0, 1.0)
        nvp[c, k, 0], nvp[c, k, 1] = components(gx * flip, gy * flip, gz * flip, qx, qy)
    nvt = np.zeros((nV + 1, D, 2))
    interior = variational_init.interior_vertex(mesh)[:nV] == 1
    v = np.nonzero(interior)[0]
    lat_vert = np.arcsin(np.clip(vz[v] / R, -1.0, 1.0))
    lon_vert = np.arctan2(vy[v], vx[v])
    for s_ in range(D):
        e = eov[v, s_].astype(np.int64) - 1
        qx, qy, qz = to_equator(ex[e], ey[e], ez[e], lat_vert, lon_vert)
        c1x, c1y, c1z = to_equator(cx[coe[e, 0]], cy[coe[e, 0]], cz[coe[e, 0]], lat_vert, lon_vert)
        c2x, c2y, c2z = to_equator(cx[coe[e, 1]], cy[coe[e, 1]], cz[coe[e, 1]], lat_vert, lon_vert)
        wx, wy, wz = c2x - c1x, c2y - c1y, c2z - c1z
        gx, gy, gz = wy * qz - wz * qy, wz * qx - wx * qz, wx * qy - wy * qx
        flip = np.where(v == a[e], -1.0, 1.0)
        nvt[v, s_, 0], nvt[v, s_, 1] = components(gx * flip, gy * flip, gz * flip, qx, qy)
    lat_c = np.zeros(nC + 1)
    lat_v = np.zeros(nV + 1)
    lat_c[:nC] = np.arcsin(np.clip(cz[:nC] / R, -1.0, 1.0))
    lat_v[:nV][interior] = np.arcsin(np.clip(vz[:nV][interior] / R, -1.0, 1.0))
    return dict(verticesOnEdge=voe, edgesOnVertex=eov, normalVectorPolygon=nvp, normalVectorTriangle=nvt,
                latCellRotated=lat_c, latVertexRotated=lat_v)


def weak_fields(mesh):
    """dict(verticesOnEdge, edgesOnVertex, normalVectorPolygon (nCells+1, maxEdges, 2), normalVectorTriangle
    (nVertices+1, vertexDegree, 2), latCellRotated, latVertexRotated) in Registry layouts."""
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    voe, eov = edge_connectivity(mesh)
    on_sphere = bool(mesh.on_a_sphere)
    if not on_sphere:
        return _planar_fields(mesh, voe, eov)
    return _sphere_fields(mesh, voe, eov)
