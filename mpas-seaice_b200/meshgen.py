"""Offline mesh generators in MPAS conventions (the "meshes generated offline" of the north star).

Produces exactly the Registry ``mesh`` fields the EVP path reads
(reference: src/Registry.xml:2251-2367): nEdgesOnCell, verticesOnCell, cellsOnVertex, cellsOnCell,
edgesOnCell, x/y/zCell, x/y/zVertex, latVertex, areaCell, areaTriangle, kiteAreasOnVertex, dvEdge,
dcEdge, fVertex.

Conventions (MPAS in-memory, cf. src/shared/mpas_seaice_initialize.F:214-234):
  * index VALUES are 1-based; an invalid neighbour points at the junk slot ``n+1``;
  * every array has one extra (junk) element at the end: ``nEdgesOnCell[nCells] = 0``,
    ``areaCell[nCells] = -1e34``;
  * 2-D arrays are stored C-order with shape ``(nCells+1, maxEdges)`` which is byte-identical to the
    Fortran column-major ``(maxEdges, nCells+1)``;
  * polygons are counter-clockwise; edge ``s`` of a cell joins vertex ``s`` and ``s+1``
    (the assumption made by src/shared/mpas_seaice_velocity_solver_pwl.F:161-176);
    ``cellsOnCell(s, c)`` is the cell across edge ``s``.

Three families:
  planar_hex(nx, ny, dc)      -- the culled periodic_hex square of the reference test cases
                                 (testing_and_setup/testcases/square/*/create_grids.py)
  planar_quad(nx, ny, dc)     -- the quad square (create_grids.py:66-170)
  icosphere(level, radius)    -- icosahedral Voronoi sphere, 10*4**level + 2 cells
                                 (QU240 ~ level 5, QU60 ~ level 7, QU7.5 ~ level 10)

Everything is vectorised numpy so the 10.5 M-cell sphere is generated in about a minute.
"""
from __future__ import annotations

import numpy as np

EARTH_RADIUS = 6371229.0  # reference: src/Registry.xml:447 (config_earth_radius)
OMEGA = 7.29212e-5        # reference: src/shared/mpas_seaice_constants.F (seaiceOmega)
JUNK_AREA = -1.0e34       # reference: src/shared/mpas_seaice_initialize.F:232-234


class Mesh(dict):
    """dict with attribute access; all arrays carry the MPAS junk slot."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


# ----------------------------------------------------------------------------------------------
# helpers shared by the planar generators
# ----------------------------------------------------------------------------------------------

def _finish_planar(xc, yc, vx, vy, voc, n_on_cell, max_edges, vertex_degree, fvertex_value):
    """Build all derived connectivity + geometry for a planar polygon mesh.

    voc: (nCells, maxEdges) 0-based vertex ids, CCW, -1 padded.
    """
    nC = xc.shape[0]
    nV = vx.shape[0]
    M = max_edges
    D = vertex_degree

    valid = voc >= 0
    # ---- cellsOnVertex, CCW around the vertex --------------------------------------------
    cell_of = np.repeat(np.arange(nC), M).reshape(nC, M)[valid]
    vert_of = voc[valid]
    ang = np.arctan2(yc[cell_of] - vy[vert_of], xc[cell_of] - vx[vert_of])
    order = np.lexsort((ang, vert_of))
    vert_s = vert_of[order]
    cell_s = cell_of[order]
    start = np.searchsorted(vert_s, np.arange(nV))
    count = np.searchsorted(vert_s, np.arange(nV), side="right") - start
    cov = np.full((nV, D), -1, dtype=np.int64)
    for s in range(D):
        has = count > s
        cov[has, s] = cell_s[start[has] + s]
    # for boundary vertices with a gap in the fan, rotate so that the fan is contiguous CCW:
    # (start after the largest angular gap).  Keeps MPAS-like "valid cells first is NOT required".
    if D >= 3:
        ang_s = ang[order]
        for v in np.nonzero((count < D) & (count > 1))[0]:
            a = ang_s[start[v]:start[v] + count[v]]
            gaps = np.diff(np.concatenate([a, a[:1] + 2 * np.pi]))
            k = int(np.argmax(gaps)) + 1
            if k < count[v]:
                cov[v, :count[v]] = np.roll(cov[v, :count[v]], -k)

    # ---- edges ---------------------------------------------------------------------------
    nxt = np.empty_like(voc)
    for c_n in np.unique(n_on_cell):
        rows = n_on_cell == c_n
        sub = voc[rows]
        nx_ = sub.copy()
        nx_[:, :c_n] = np.roll(sub[:, :c_n], -1, axis=1)
        nxt[rows] = nx_
    v1 = voc[valid]
    v2 = nxt[valid]
    lo = np.minimum(v1, v2)
    hi = np.maximum(v1, v2)
    key = lo * np.int64(nV) + hi
    ukey, inv = np.unique(key, return_inverse=True)
    nE = ukey.shape[0]
    eoc = np.full((nC, M), -1, dtype=np.int64)
    eoc[valid] = inv
    ev1 = (ukey // nV).astype(np.int64)
    ev2 = (ukey % nV).astype(np.int64)
    dv_edge = np.hypot(vx[ev1] - vx[ev2], vy[ev1] - vy[ev2])
    # cells on edge (up to 2)
    slot_cell = cell_of
    order_e = np.argsort(inv, kind="stable")
    inv_s = inv[order_e]
    cell_e = slot_cell[order_e]
    first = np.searchsorted(inv_s, np.arange(nE))
    cnt_e = np.searchsorted(inv_s, np.arange(nE), side="right") - first
    coe = np.full((nE, 2), -1, dtype=np.int64)
    coe[:, 0] = cell_e[first]
    two = cnt_e > 1
    coe[two, 1] = cell_e[first[two] + 1]
    # cellsOnCell: other cell on edge s
    coc = np.full((nC, M), -1, dtype=np.int64)
    e_flat = eoc[valid]
    other = np.where(coe[e_flat, 0] == cell_of, coe[e_flat, 1], coe[e_flat, 0])
    coc[valid] = other
    dc_edge = np.zeros(nE)
    dc_edge[two] = np.hypot(xc[coe[two, 0]] - xc[coe[two, 1]], yc[coe[two, 0]] - yc[coe[two, 1]])

    # ---- areas ---------------------------------------------------------------------------
    # shoelace for cells
    area_cell = np.zeros(nC)
    for s in range(M):
        ok = valid[:, s]
        a = voc[ok, s]
        b = nxt[ok, s]
        area_cell[ok] += 0.5 * ((vx[a] - xc[ok]) * (vy[b] - yc[ok]) - (vx[b] - xc[ok]) * (vy[a] - yc[ok]))
    # kites: for vertex v and cell c (slot j in cell): quad (centre, mid(prev edge), v, mid(next edge))
    kite = np.zeros((nV, D))
    prv = np.empty_like(voc)
    for c_n in np.unique(n_on_cell):
        rows = n_on_cell == c_n
        sub = voc[rows]
        pv = sub.copy()
        pv[:, :c_n] = np.roll(sub[:, :c_n], 1, axis=1)
        prv[rows] = pv
    for s in range(D):
        ok = cov[:, s] >= 0
        vs = np.nonzero(ok)[0]
        cs = cov[vs, s]
        # slot of v in cell
        j = np.argmax(voc[cs] == vs[:, None], axis=1)
        pvv = prv[cs, j]
        nvv = nxt[cs, j]
        px = np.stack([xc[cs], 0.5 * (vx[pvv] + vx[vs]), vx[vs], 0.5 * (vx[nvv] + vx[vs])], axis=1)
        py = np.stack([yc[cs], 0.5 * (vy[pvv] + vy[vs]), vy[vs], 0.5 * (vy[nvv] + vy[vs])], axis=1)
        kite[vs, s] = 0.5 * np.abs(np.sum(px * np.roll(py, -1, axis=1) - np.roll(px, -1, axis=1) * py, axis=1))
    area_tri = kite.sum(axis=1)

    m = Mesh()
    m.on_a_sphere = False
    m.sphere_radius = 1.0
    m.nCells, m.nVertices, m.nEdges = nC, nV, nE
    m.maxEdges, m.vertexDegree = M, D

    def pad1(a, junk, dtype):
        out = np.empty(a.shape[0] + 1, dtype=dtype)
        out[:-1] = a
        out[-1] = junk
        return out

    def pad_idx(a, n_target):
        out = np.empty((a.shape[0] + 1, a.shape[1]), dtype=np.int32)
        out[:-1] = np.where(a >= 0, a + 1, n_target + 1)
        out[-1] = n_target + 1
        return out

    m.nEdgesOnCell = pad1(n_on_cell, 0, np.int32)
    m.verticesOnCell = pad_idx(voc, nV)
    m.edgesOnCell = pad_idx(eoc, nE)
    m.cellsOnCell = pad_idx(coc, nC)
    m.cellsOnVertex = pad_idx(cov, nC)
    m.cellsOnEdge = pad_idx(coe, nC)
    m.xCell, m.yCell, m.zCell = pad1(xc, 0.0, np.float64), pad1(yc, 0.0, np.float64), np.zeros(nC + 1)
    m.xVertex, m.yVertex, m.zVertex = pad1(vx, 0.0, np.float64), pad1(vy, 0.0, np.float64), np.zeros(nV + 1)
    m.latCell, m.lonCell = np.zeros(nC + 1), np.zeros(nC + 1)
    m.latVertex, m.lonVertex = np.zeros(nV + 1), np.zeros(nV + 1)
    m.areaCell = pad1(area_cell, JUNK_AREA, np.float64)
    m.areaTriangle = pad1(area_tri, 0.0, np.float64)
    m.kiteAreasOnVertex = np.vstack([kite, np.zeros((1, D))])
    m.dvEdge = pad1(dv_edge, 0.0, np.float64)
    m.dcEdge = pad1(dc_edge, 0.0, np.float64)
    m.fVertex = pad1(np.full(nV, fvertex_value), 0.0, np.float64)
    return m


def planar_hex(nx: int, ny: int, dc: float, fvertex: float = 1.46e-4) -> Mesh:
    """nx x ny pointy-top hexagons, row j shifted by dc/2 for odd j: what MPAS-Tools ``periodic_grid``
    followed by culling of the periodic boundary rows leaves
    (reference recipe: testing_and_setup/testcases/square/operators_strain_stress_divergence/create_grids.py:9-64).
    fVertex default = the square test case value (square_quadhex/create_ics.py:92-93)."""
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ii = ii.ravel()
    jj = jj.ravel()
    # integer lattice: x in units of dc/2, y in units of dc/(2 sqrt 3)
    cx = 2 * ii + (jj % 2) + 1
    cy = 3 * jj + 2
    off = np.array([[1, 1], [0, 2], [-1, 1], [-1, -1], [0, -2], [1, -1]])  # 30,90,...,330 degrees: CCW
    kx = cx[:, None] + off[None, :, 0]
    ky = cy[:, None] + off[None, :, 1]
    W = 2 * nx + 4
    key = ky.astype(np.int64) * W + kx
    ukey, inv = np.unique(key.ravel(), return_inverse=True)
    voc = inv.reshape(-1, 6).astype(np.int64)
    ux = 0.5 * dc
    uy = dc / (2.0 * np.sqrt(3.0))
    vx = (ukey % W).astype(np.float64) * ux
    vy = (ukey // W).astype(np.float64) * uy
    xc = cx.astype(np.float64) * ux
    yc = cy.astype(np.float64) * uy
    n_on = np.full(nx * ny, 6, dtype=np.int64)
    m = _finish_planar(xc, yc, vx, vy, voc, n_on, 6, 3, fvertex)
    m.nx, m.ny, m.dc = nx, ny, dc
    m.Lx, m.Ly = nx * dc, ny * dc
    m.kind = "planar_hex"
    return m


def planar_quad(nx: int, ny: int, dc: float, fvertex: float = 1.46e-4) -> Mesh:
    """nx x ny squares (reference recipe: create_grids.py:66-170; vertexDegree = 4)."""
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ii = ii.ravel()
    jj = jj.ravel()
    xc = (ii + 0.5) * dc
    yc = (jj + 0.5) * dc
    vi, vj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    vx = vi.ravel() * float(dc)
    vy = vj.ravel() * float(dc)
    vid = lambda i, j: i + j * (nx + 1)
    voc = np.stack([vid(ii, jj), vid(ii + 1, jj), vid(ii + 1, jj + 1), vid(ii, jj + 1)], axis=1).astype(np.int64)
    n_on = np.full(nx * ny, 4, dtype=np.int64)
    m = _finish_planar(xc, yc, vx, vy, voc, n_on, 4, 4, fvertex)
    m.nx, m.ny, m.dc = nx, ny, dc
    m.Lx, m.Ly = nx * dc, ny * dc
    m.kind = "planar_quad"
    return m


# ----------------------------------------------------------------------------------------------
# icosahedral Voronoi sphere
# ----------------------------------------------------------------------------------------------

def _icosahedron():
    t = (1.0 + np.sqrt(5.0)) / 2.0
    p = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                  [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    p /= np.linalg.norm(p, axis=1)[:, None]
    # tilt slightly so that no point sits exactly on a pole / on z = 0 of the rotated grid
    a, b = 0.3711, 0.2137
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    p = p @ (Rx @ Ry).T
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    # make every face CCW seen from outside
    n = np.cross(p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]])
    flip = np.einsum("ij,ij->i", n, p[f[:, 0]]) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    return p, f


def _subdivide(p, f):
    nP = p.shape[0]
    a, b, c = f[:, 0], f[:, 1], f[:, 2]
    e = np.concatenate([np.stack([a, b], 1), np.stack([b, c], 1), np.stack([c, a], 1)], axis=0)
    lo = e.min(axis=1)
    hi = e.max(axis=1)
    key = lo * np.int64(nP) + hi
    ukey, inv = np.unique(key, return_inverse=True)
    mid = _normalize(p[ukey // nP] + p[ukey % nP])
    p2 = np.concatenate([p, mid], axis=0)
    nF = f.shape[0]
    mab = nP + inv[:nF]
    mbc = nP + inv[nF:2 * nF]
    mca = nP + inv[2 * nF:]
    # children stored contiguously (4t..4t+3): triangle numbering is a quadtree order => locality
    f2 = np.empty((nF, 4, 3), dtype=np.int64)
    f2[:, 0] = np.stack([a, mab, mca], 1)
    f2[:, 1] = np.stack([mab, b, mbc], 1)
    f2[:, 2] = np.stack([mab, mbc, mca], 1)
    f2[:, 3] = np.stack([mca, mbc, c], 1)
    return p2, f2.reshape(-1, 3)


def _cross(a, b):
    """Row-wise cross product (much faster than np.cross for (n, 3) arrays)."""
    out = np.empty_like(a)
    out[:, 0] = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    out[:, 1] = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    out[:, 2] = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return out


def _dot(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def _normalize(a):
    a /= np.sqrt(_dot(a, a))[:, None]
    return a


def _first_index(keys, n):
    """first_index[k] = smallest i with keys[i] == k (keys in 0..n-1, every k present)."""
    out = np.full(n, keys.shape[0], dtype=np.int64)
    idx = np.arange(keys.shape[0] - 1, -1, -1, dtype=np.int64)
    out[keys[::-1]] = idx          # repeated indices: the last assignment (= smallest i) wins
    return out


def _sph_tri_area(a, b, c):
    """Spherical triangle area on the unit sphere (Van Oosterom & Strackee), vectorised."""
    num = np.abs(_dot(a, _cross(b, c)))
    den = 1.0 + _dot(a, b) + _dot(b, c) + _dot(c, a)
    return 2.0 * np.arctan2(num, den)


def icosphere(level: int, radius: float = EARTH_RADIUS) -> Mesh:
    """Voronoi dual of the level-times bisected icosahedron: 10*4**level+2 cells (12 pentagons),
    20*4**level vertices, vertexDegree 3, maxEdges 6.  Cells and vertices are numbered along the
    quadtree order of the triangulation (good L2 locality)."""
    p, f = _icosahedron()
    for _ in range(level):
        p, f = _subdivide(p, f)
    nC0 = p.shape[0]
    nV = f.shape[0]
    # renumber cells by first incident triangle (triangles are in quadtree order)
    first_tri = _first_index(f.reshape(-1), nC0) // 3
    perm = np.argsort(first_tri, kind="stable")      # new -> old
    inv_perm = np.empty(nC0, dtype=np.int64)
    inv_perm[perm] = np.arange(nC0)
    p = p[perm]
    f = inv_perm[f]
    nC = nC0

    # MPAS vertices = triangle circumcentres on the sphere
    a, b, c = p[f[:, 0]], p[f[:, 1]], p[f[:, 2]]
    vpos = _normalize(_cross(b - a, c - a))
    del a, b, c

    # rings: record r = (triangle t, corner k): cell = f[t,k], next vertex = f[t,k+1], prev = f[t,k+2]
    # going CCW around a cell, the triangle after (cell, nxt, prv) is the one whose "nxt" equals prv.
    t_id = np.repeat(np.arange(nV), 3)
    cell = f.reshape(-1)
    nxt = np.roll(f, -1, axis=1).reshape(-1)
    prv = np.roll(f, -2, axis=1).reshape(-1)
    key = cell * np.int64(nC) + nxt
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    link = np.searchsorted(key_s, cell * np.int64(nC) + prv)
    nxt_rec = order[link]                               # record index of the next triangle around cell
    # start record per cell: the record with the smallest triangle id
    start = _first_index(cell, nC)
    ring = np.empty((nC, 6), dtype=np.int64)
    ringn = np.empty((nC, 6), dtype=np.int64)           # neighbour cell across edge s (between tri s and s+1)
    cur = start.copy()
    n_on = np.full(nC, 6, dtype=np.int64)
    for s in range(6):
        ring[:, s] = t_id[cur]
        ringn[:, s] = prv[cur]
        cur = nxt_rec[cur]
        if s == 4:
            n_on[cur == start] = 5
    pent = n_on == 5
    ring[pent, 5] = -1
    ringn[pent, 5] = -1
    voc = ring
    coc = ringn
    # cellsOnVertex: corners of the triangle, CCW
    cov = f.copy()

    # edges: one per (cell, neighbour) pair
    valid = voc >= 0
    cell_of = np.repeat(np.arange(nC), 6).reshape(nC, 6)[valid]
    nb = coc[valid]
    lo = np.minimum(cell_of, nb)
    hi = np.maximum(cell_of, nb)
    ukey, inv = np.unique(lo * np.int64(nC) + hi, return_inverse=True)
    nE = ukey.shape[0]
    eoc = np.full((nC, 6), -1, dtype=np.int64)
    eoc[valid] = inv
    ec1 = ukey // nC
    ec2 = ukey % nC
    nxt_v = voc.copy()
    for n_ in (5, 6):
        rows = n_on == n_
        sub = voc[rows]
        sub2 = sub.copy()
        sub2[:, :n_] = np.roll(sub[:, :n_], -1, axis=1)
        nxt_v[rows] = sub2
    # edge vertices from the first cell listing it
    ev1 = np.empty(nE, dtype=np.int64)
    ev2 = np.empty(nE, dtype=np.int64)
    ev1[inv] = voc[valid]
    ev2[inv] = nxt_v[valid]

    def arc(u, v):
        w = _cross(u, v)
        return np.arctan2(np.sqrt(_dot(w, w)), _dot(u, v))

    dv_edge = arc(vpos[ev1], vpos[ev2]) * radius
    dc_edge = arc(p[ec1], p[ec2]) * radius

    # areas (unit sphere, then scaled)
    area_cell = np.zeros(nC)
    for s in range(6):
        ok = valid[:, s]
        area_cell[ok] += _sph_tri_area(p[ok], vpos[voc[ok, s]], vpos[nxt_v[ok, s]])
    # kites: (vertex v, cell corner k): cell centre, mid of edge to previous corner cell, v, mid to next
    kite = np.zeros((nV, 3))
    for k in range(3):
        cc = p[f[:, k]]
        cn = p[f[:, (k + 1) % 3]]
        cp = p[f[:, (k + 2) % 3]]
        m1 = _normalize(cc + cn)
        m2 = _normalize(cc + cp)
        kite[:, k] = _sph_tri_area(cc, m1, vpos) + _sph_tri_area(cc, vpos, m2)
    kite *= radius * radius
    area_cell *= radius * radius

    m = Mesh()
    m.on_a_sphere = True
    m.sphere_radius = float(radius)
    m.nCells, m.nVertices, m.nEdges = nC, nV, nE
    m.maxEdges, m.vertexDegree = 6, 3
    m.level = level
    m.kind = "icosphere"

    def pad1(arr, junk, dtype):
        out = np.empty(arr.shape[0] + 1, dtype=dtype)
        out[:-1] = arr
        out[-1] = junk
        return out

    def pad_idx(arr, n_target):
        out = np.empty((arr.shape[0] + 1, arr.shape[1]), dtype=np.int32)
        out[:-1] = np.where(arr >= 0, arr + 1, n_target + 1)
        out[-1] = n_target + 1
        return out

    m.nEdgesOnCell = pad1(n_on, 0, np.int32)
    m.verticesOnCell = pad_idx(voc, nV)
    m.edgesOnCell = pad_idx(eoc, nE)
    m.cellsOnCell = pad_idx(coc, nC)
    m.cellsOnVertex = pad_idx(cov, nC)
    m.cellsOnEdge = pad_idx(np.stack([ec1, ec2], 1), nC)
    pc = p * radius
    pv = vpos * radius
    m.xCell, m.yCell, m.zCell = (pad1(pc[:, i], 0.0, np.float64) for i in range(3))
    m.xVertex, m.yVertex, m.zVertex = (pad1(pv[:, i], 0.0, np.float64) for i in range(3))
    m.latCell = pad1(np.arcsin(np.clip(p[:, 2], -1, 1)), 0.0, np.float64)
    m.lonCell = pad1(np.arctan2(p[:, 1], p[:, 0]), 0.0, np.float64)
    m.latVertex = pad1(np.arcsin(np.clip(vpos[:, 2], -1, 1)), 0.0, np.float64)
    m.lonVertex = pad1(np.arctan2(vpos[:, 1], vpos[:, 0]), 0.0, np.float64)
    m.areaCell = pad1(area_cell, JUNK_AREA, np.float64)
    m.kiteAreasOnVertex = np.vstack([kite, np.zeros((1, 3))])
    # areaTriangle = sum of kites, as recomputed by the reference (src/shared/mpas_seaice_mesh.F:234-240)
    at = np.zeros(nV)
    for k in range(3):
        at = at + kite[:, k]
    m.areaTriangle = pad1(at, 0.0, np.float64)
    m.dvEdge = pad1(dv_edge, 0.0, np.float64)
    m.dcEdge = pad1(dc_edge, 0.0, np.float64)
    m.fVertex = pad1(2.0 * OMEGA * np.sin(m.latVertex[:-1]), 0.0, np.float64)
    return m


def latlon_band(nlon: int, nlat: int, lat_max_deg: float = 60.0, radius: float = EARTH_RADIUS) -> Mesh:
    """A quadrilateral mesh on the sphere: nlon x nlat cells between -lat_max and +lat_max, periodic in longitude,
    every interior vertex of degree 4 -- the kind of mesh the reference's "spherical quad" notes refer to
    (src/shared/mpas_seaice_advection_incremental_remap.F:105-119), here without poles.  Only the fields the
    transport scheme reads (connectivity, coordinates, areaCell, dcEdge, dvEdge)."""
    dlon, lat0 = 2.0 * np.pi / nlon, -np.deg2rad(lat_max_deg)
    dlat = -2.0 * lat0 / nlat
    nC, nV = nlon * nlat, nlon * (nlat + 1)
    nH = nlon * (nlat + 1)                      # edges along parallels: h(i, j) joins v(i, j) and v(i+1, j)
    nE = nH + nlon * nlat                       # edges along meridians: w(i, j) joins v(i, j) and v(i, j+1)
    ii, jj = np.meshgrid(np.arange(nlon), np.arange(nlat), indexing="xy")
    ii, jj = ii.ravel(), jj.ravel()             # cell c = j * nlon + i
    ip = (ii + 1) % nlon
    im = (ii - 1) % nlon

    def vid(i, j):
        return j * nlon + i

    def unit(lon, lat):
        return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)

    def arc(a, b):
        return radius * np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), np.sum(a * b, axis=-1))
    voc = np.stack([vid(ii, jj), vid(ip, jj), vid(ip, jj + 1), vid(ii, jj + 1)], axis=1)            # counter-clockwise
    eoc = np.stack([jj * nlon + ii, nH + jj * nlon + ip, (jj + 1) * nlon + ii, nH + jj * nlon + ii], axis=1)
    south = np.where(jj > 0, (jj - 1) * nlon + ii, -1)
    north = np.where(jj < nlat - 1, (jj + 1) * nlon + ii, -1)
    coc = np.stack([south, jj * nlon + ip, north, jj * nlon + im], axis=1)
    vi, vj = np.meshgrid(np.arange(nlon), np.arange(nlat + 1), indexing="xy")
    vi, vj = vi.ravel(), vj.ravel()
    vim = (vi - 1) % nlon
    below, above = vj > 0, vj < nlat
    cov = np.stack([np.where(below, (vj - 1) * nlon + vim, -1), np.where(below, (vj - 1) * nlon + vi, -1),
                    np.where(above, vj * nlon + vi, -1), np.where(above, vj * nlon + vim, -1)], axis=1)
    # cells of every edge: parallels have (south, north), meridians (west, east)
    hi, hj = vi, vj
    coe_h = np.stack([np.where(hj > 0, (hj - 1) * nlon + hi, -1), np.where(hj < nlat, hj * nlon + hi, -1)], axis=1)
    coe_w = np.stack([jj * nlon + im, jj * nlon + ii], axis=1)
    coe = np.concatenate([coe_h, coe_w])
    lon_v, lat_v = vi * dlon, lat0 + vj * dlat
    lon_c, lat_c = (ii + 0.5) * dlon, lat0 + (jj + 0.5) * dlat
    pv, pc = unit(lon_v, lat_v), unit(lon_c, lat_c)
    area = radius * radius * dlon * (np.sin(lat0 + (jj + 1) * dlat) - np.sin(lat0 + jj * dlat))
    ev = np.concatenate([np.stack([vid(hi, hj), vid((hi + 1) % nlon, hj)], axis=1), np.stack([vid(ii, jj), vid(ii, jj + 1)], axis=1)])
    dv = arc(pv[ev[:, 0]], pv[ev[:, 1]])
    mid = pv[ev[:, 0]] + pv[ev[:, 1]]
    mid /= np.linalg.norm(mid, axis=1)[:, None]
    both = (coe[:, 0] >= 0) & (coe[:, 1] >= 0)
    one = np.where(coe[:, 0] >= 0, coe[:, 0], coe[:, 1])
    dc = np.where(both, arc(pc[np.maximum(coe[:, 0], 0)], pc[np.maximum(coe[:, 1], 0)]), 2.0 * arc(pc[one], mid))

    m = Mesh()
    m.on_a_sphere, m.sphere_radius, m.kind = True, radius, "latlon_band"
    m.nCells, m.nVertices, m.nEdges, m.maxEdges, m.vertexDegree = nC, nV, nE, 4, 4

    def pad1(a, junk, dtype):
        out = np.empty(a.shape[0] + 1, dtype=dtype)
        out[:-1] = a
        out[-1] = junk
        return out

    def pad_idx(a, n_target):
        out = np.empty((a.shape[0] + 1, a.shape[1]), dtype=np.int32)
        out[:-1] = np.where(a >= 0, a + 1, n_target + 1)
        out[-1] = n_target + 1
        return out
    m.nEdgesOnCell = pad1(np.full(nC, 4), 0, np.int32)
    m.verticesOnCell, m.edgesOnCell, m.cellsOnCell = pad_idx(voc, nV), pad_idx(eoc, nE), pad_idx(coc, nC)
    m.cellsOnVertex, m.cellsOnEdge = pad_idx(cov, nC), pad_idx(coe, nC)
    m.xCell, m.yCell, m.zCell = (pad1(radius * pc[:, k], 0.0, np.float64) for k in range(3))
    m.xVertex, m.yVertex, m.zVertex = (pad1(radius * pv[:, k], 0.0, np.float64) for k in range(3))
    m.latCell, m.lonCell = pad1(lat_c, 0.0, np.float64), pad1(lon_c, 0.0, np.float64)
    m.latVertex, m.lonVertex = pad1(lat_v, 0.0, np.float64), pad1(lon_v, 0.0, np.float64)
    m.areaCell = pad1(area, JUNK_AREA, np.float64)
    m.dvEdge, m.dcEdge = pad1(dv, 0.0, np.float64), pad1(dc, 0.0, np.float64)
    return m


def check_mesh(m: Mesh) -> None:
    """Structural invariants every generator must satisfy (used by tests)."""
    nC, nV, M, D = m.nCells, m.nVertices, m.maxEdges, m.vertexDegree
    voc = m.verticesOnCell
    cov = m.cellsOnVertex
    assert voc.shape == (nC + 1, M) and cov.shape == (nV + 1, D)
    assert m.nEdgesOnCell[nC] == 0
    n = m.nEdgesOnCell[:nC]
    for s in range(M):
        ok = n > s
        assert np.all((voc[:nC][ok, s] >= 1) & (voc[:nC][ok, s] <= nV))
    # every (cell, vertex) incidence appears in cellsOnVertex
    for s in range(M):
        ok = np.nonzero(n > s)[0]
        v = voc[ok, s] - 1
        assert np.all(np.any(cov[v] == (ok + 1)[:, None], axis=1))
    # CCW
    if not m.on_a_sphere:
        x = m.xVertex
        y = m.yVertex
        area = np.zeros(nC)
        for s in range(M):
            ok = n > s
            a = voc[:nC][ok, s] - 1
            b = voc[np.nonzero(ok)[0], (s + 1) % n[ok]] - 1
            area[ok] += 0.5 * (x[a] * y[b] - x[b] * y[a])
        assert np.all(area > 0)
