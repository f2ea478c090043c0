"""Mesh-file quantities the incremental-remapping transport reads, for non-Fortran hosts.

In MPAS-Seaice ``verticesOnEdge``, ``edgesOnVertex`` and ``x/y/zEdge`` come from the mesh file and
``coeffs_reconstruct`` from the MPAS framework (``mpas_init_reconstruct`` in mpas_vector_reconstruction.F, called
at src/shared/mpas_seaice_advection_incremental_remap.F:744-746 -- an un-vendored dependency of the reference,
src/Makefile:6-7).  The generators of meshgen.py emit none of them, so this module derives them:

* ``verticesOnEdge`` oriented the MPAS way -- ``cellsOnEdge(1)`` lies to the left of V1 -> V2, which
  get_geometry_incremental_remap checks (incremental_remap.F:1290-1330);
* ``x/y/zEdge``: the point half way between the two cell centres (on the sphere: on the great-circle arc), which on
  a Voronoi mesh lies on the edge; for an edge with one cell the mid-point of its vertices;
* ``coeffs_reconstruct``: the published algorithm of the framework routine -- radial basis functions
  1/sqrt(1 + r^2/alpha^2) on the edge points of a cell, normal components as data, constants in the tangent plane
  reproduced exactly, alpha = mean over the edges of half the centre-to-edge distance.  The framework source is not
  in this image, so this is a restatement from its documentation: "parity unpinned"; a Fortran host passes the
  framework's own array.
"""
from __future__ import annotations

import numpy as np

from . import weakmesh


def _unit(v):
    n = np.sqrt(np.sum(v * v, axis=-1, keepdims=True))
    return v / np.where(n > 0, n, 1.0)


def edge_points(mesh):
    """x/y/zEdge, (nEdges+1,) each."""
    nC, nE = mesh.nCells, mesh.nEdges
    voe, _ = weakmesh.edge_connectivity(mesh)
    coe = mesh.cellsOnEdge[:nE].astype(np.int64)
    pc = np.stack([mesh.xCell, mesh.yCell, mesh.zCell], axis=1)
    pv = np.stack([mesh.xVertex, mesh.yVertex, mesh.zVertex], axis=1)
    both = (coe[:, 0] <= nC) & (coe[:, 1] <= nC)
    mid_c = 0.5 * (pc[np.minimum(coe[:, 0], nC + 1) - 1] + pc[np.minimum(coe[:, 1], nC + 1) - 1])
    mid_v = 0.5 * (pv[voe[:nE, 0] - 1] + pv[voe[:nE, 1] - 1])
    pe = np.where(both[:, None], mid_c, mid_v)
    if mesh.on_a_sphere:
        pe = _unit(pe) * mesh.sphere_radius
    out = np.zeros((nE + 1, 3))
    out[:nE] = pe
    return out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()


def oriented_vertices_on_edge(mesh):
    """verticesOnEdge (nEdges+1, 2) with cellsOnEdge(1) to the left of V1 -> V2, and edgesOnVertex."""
    nC, nE = mesh.nCells, mesh.nEdges
    voe, eov = weakmesh.edge_connectivity(mesh)
    voe = voe.copy()
    coe = mesh.cellsOnEdge[:nE].astype(np.int64)
    pc = np.stack([mesh.xCell, mesh.yCell, mesh.zCell], axis=1)
    pv = np.stack([mesh.xVertex, mesh.yVertex, mesh.zVertex], axis=1)
    v1, v2 = pv[voe[:nE, 0] - 1], pv[voe[:nE, 1] - 1]
    has1 = coe[:, 0] <= nC
    ref = np.where(has1[:, None], pc[np.minimum(coe[:, 0], nC + 1) - 1], pc[np.minimum(coe[:, 1], nC + 1) - 1])
    up = _unit(0.5 * (v1 + v2)) if mesh.on_a_sphere else np.array([0.0, 0.0, 1.0])[None, :]
    side = np.sum(np.cross(v2 - v1, ref - v1) * up, axis=1)
    side = np.where(has1, side, -side)       # no first cell: keep the second cell to the right
    flip = side < 0
    voe[:nE][flip] = voe[:nE][flip][:, ::-1]
    return voe, eov


def reconstruction_coefficients(mesh, x_edge, y_edge, z_edge, chunk=262144):
    """coeffs_reconstruct (nCells+1, maxEdges, 3): gradient (or any tangent vector) at a cell centre =
    sum over the cell's edges of coeffs * (normal component at the edge, positive from cellsOnEdge(1) to (2))."""
    nC, nE, M = mesh.nCells, mesh.nEdges, mesh.maxEdges
    pc = np.stack([mesh.xCell, mesh.yCell, mesh.zCell], axis=1)[:nC]
    pe = np.stack([x_edge, y_edge, z_edge], axis=1)
    coe = mesh.cellsOnEdge[:nE].astype(np.int64)
    pc_all = np.stack([mesh.xCell, mesh.yCell, mesh.zCell], axis=1)
    # edge normals: from cell 1 to cell 2; at a boundary from the one cell to the edge point
    n_e = np.zeros((nE + 1, 3))
    c1ok, c2ok = coe[:, 0] <= nC, coe[:, 1] <= nC
    d = pc_all[np.minimum(coe[:, 1], nC + 1) - 1] - pc_all[np.minimum(coe[:, 0], nC + 1) - 1]
    d = np.where((c1ok & ~c2ok)[:, None], pe[:nE] - pc_all[np.minimum(coe[:, 0], nC + 1) - 1], d)
    d = np.where((~c1ok & c2ok)[:, None], pe[:nE] - pc_all[np.minimum(coe[:, 1], nC + 1) - 1], d)
    n_e[:nE] = _unit(d)
    # tangent plane of every cell
    if mesh.on_a_sphere:
        r = _unit(pc)
        helper = np.where(np.abs(r[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
        t1 = _unit(np.cross(helper, r))
        t2 = np.cross(r, t1)
    else:
        t1 = np.tile(np.array([[1.0, 0.0, 0.0]]), (nC, 1))
        t2 = np.tile(np.array([[0.0, 1.0, 0.0]]), (nC, 1))
    out = np.zeros((nC + 1, M, 3))
    n_on = mesh.nEdgesOnCell[:nC]
    for n in np.unique(n_on):
        if n < 1:
            continue
        group = np.nonzero(n_on == n)[0]
        for start in range(0, group.size, chunk):          # bounded memory: (chunk, n+2, n+2) systems at a time
            cells = group[start:start + chunk]
            e = mesh.edgesOnCell[cells, :n].astype(np.int64) - 1          # (B, n)
            x = pe[e]                                                     # (B, n, 3)
            nv = n_e[e]
            c = pc[cells]
            alpha = np.mean(0.5 * np.sqrt(np.sum((x - c[:, None, :]) ** 2, axis=2)), axis=1)   # (B,)
            # everything in the tangent plane
            xs = np.stack([np.sum(x * t1[cells][:, None, :], axis=2), np.sum(x * t2[cells][:, None, :], axis=2)], axis=2)
            ns = np.stack([np.sum(nv * t1[cells][:, None, :], axis=2), np.sum(nv * t2[cells][:, None, :], axis=2)], axis=2)
            cs = np.stack([np.sum(c * t1[cells], axis=1), np.sum(c * t2[cells], axis=1)], axis=1)
            r2 = np.sum((xs[:, :, None, :] - xs[:, None, :, :]) ** 2, axis=3) / (alpha ** 2)[:, None, None]
            A = np.zeros((cells.size, n + 2, n + 2))
            A[:, :n, :n] = (1.0 / np.sqrt(1.0 + r2)) * np.sum(ns[:, :, None, :] * ns[:, None, :, :], axis=3)
            A[:, :n, n:] = ns
            A[:, n:, :n] = np.transpose(ns, (0, 2, 1))
            rd = np.sum((xs - cs[:, None, :]) ** 2, axis=2) / (alpha ** 2)[:, None]
            rhs = np.zeros((cells.size, n + 2, 2))
            rhs[:, :n, :] = (1.0 / np.sqrt(1.0 + rd))[:, :, None] * ns
            rhs[:, n, 0] = 1.0
            rhs[:, n + 1, 1] = 1.0
            sol = np.linalg.solve(A, rhs)[:, :n, :]                       # (B, n, 2)
            out[cells, :n, :] = sol[:, :, 0:1] * t1[cells][:, None, :] + sol[:, :, 1:2] * t2[cells][:, None, :]
    return out


def ir_fields(mesh):
    """Everything orc_ir_init_geometry / ir_* need beyond the velocity solver's mesh arrays."""
    voe, eov = oriented_vertices_on_edge(mesh)
    xe, ye, ze = edge_points(mesh)
    return dict(verticesOnEdge=voe, edgesOnVertex=eov, xEdge=xe, yEdge=ye, zEdge=ze,
                coeffs_reconstruct=reconstruction_coefficients(mesh, xe, ye, ze))
