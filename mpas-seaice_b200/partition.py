"""MPAS-style graph decomposition of a mesh into one block per rank (= per GPU), halo layers, local
renumbering and the send/recv lists of the per-subcycle uVelocity/vVelocity halo exchange
(reference call site: src/shared/mpas_seaice_velocity_solver.F:2543-2584, exchange group built at
:259-349; decomposition file named by config_block_decomp_file_prefix, src/Registry.xml:369;
config_num_halos, src/Registry.xml:337).

What the reference leaves to code that is NOT under /root/reference (MPAS framework block creator,
METIS ``gpmetis``) is restated here from the published MPAS conventions and is therefore
"parity unpinned" (SURVEY.md section 8c):

  * the CELL graph (cellsOnCell) is partitioned; ``graph.info`` / ``graph.info.part.N`` files in the
    METIS text format can be written and read, so a real gpmetis partition drops in;
  * a block = owned cells (ascending global id) followed by halo layer 1, 2, ... (edge adjacency);
  * a vertex is owned by the owner of the first valid cell of cellsOnVertex(:, v);
  * local vertices = owned vertices (ascending global id) followed by the halo vertices of all local
    cells, grouped by owning rank and ascending global id inside a group -- so what a neighbour
    sends lands in ONE contiguous slice of the local arrays (no unpack pass on the device);
  * slot order inside verticesOnCell / cellsOnVertex / cellsOnCell is kept, only the index values are
    renumbered: every floating-point sum of the subcycle runs in the same order as on one rank,
    which is what makes owned results bit-identical for every rank count (the reference's
    "parallelism" test policy, testing_and_setup/testing/tests/parallelism.py:75-85).

Halo depth: the reference allocates config_num_halos = 2 cell layers.  The subcycle itself reads, for
an owned vertex, only the vertexDegree cells around it; with vertexDegree = 3 those are edge
neighbours of an owned cell, so ONE layer closes the stencil (default here for hex / Voronoi meshes),
with vertexDegree = 4 (quads) the diagonal cell needs the second layer (default 2).  ``build_block``
verifies the closure and raises if ``n_halos`` is too small.
"""
from __future__ import annotations

import numpy as np

from .meshgen import JUNK_AREA, Mesh

CELL_FIELDS_1D = ("xCell", "yCell", "zCell", "latCell", "lonCell", "areaCell")
VERTEX_FIELDS_1D = ("xVertex", "yVertex", "zVertex", "latVertex", "lonVertex", "areaTriangle", "fVertex")


# ---------------------------------------------------------------------------------------------
# partitioning the cell graph
# ---------------------------------------------------------------------------------------------

def partition_cells(mesh, n_parts: int, method: str = "auto") -> np.ndarray:
    """part[c] in 0..n_parts-1 for every global cell c (0-based), balanced to +-1 cell.

    'block' : contiguous chunks of the cell numbering.  The icosphere generator numbers cells along
              the quadtree order of the triangulation, i.e. along a space-filling curve, so chunks
              are compact patches.
    'rcb'   : recursive coordinate bisection on the cell centres (any mesh, any numbering).
    'auto'  : 'block' for the icosphere, 'rcb' otherwise."""
    nC = int(mesh.nCells)
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    if method == "auto":
        method = "block" if mesh.get("kind") == "icosphere" else "rcb"
    if method == "block":
        bounds = (np.arange(n_parts + 1, dtype=np.int64) * nC) // n_parts
        part = np.empty(nC, dtype=np.int32)
        for p in range(n_parts):
            part[bounds[p]:bounds[p + 1]] = p
        return part
    if method != "rcb":
        raise ValueError(f"unknown partition method {method!r}")
    xyz = np.stack([mesh.xCell[:nC], mesh.yCell[:nC], mesh.zCell[:nC]], axis=1)
    part = np.zeros(nC, dtype=np.int32)

    def split(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        left = np_ // 2
        k = (idx.shape[0] * left) // np_
        pts = xyz[idx]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
        # stable order along the axis, ties broken by global id => deterministic
        order = np.lexsort((idx, pts[:, axis]))
        split(np.sort(idx[order[:k]]), p0, left)
        split(np.sort(idx[order[k:]]), p0 + left, np_ - left)

    split(np.arange(nC, dtype=np.int64), 0, n_parts)
    return part


def write_graph_info(mesh, path: str) -> None:
    """METIS graph file of the cell graph, the ``graph.info`` MPAS meshes ship with: first line
    "nCells nEdges", then one line per cell with its 1-based neighbour cells."""
    nC, M = int(mesh.nCells), int(mesh.maxEdges)
    coc = mesh.cellsOnCell[:nC]
    valid = (np.arange(M)[None, :] < mesh.nEdgesOnCell[:nC, None]) & (coc >= 1) & (coc <= nC)
    n_edges = int(valid.sum()) // 2
    with open(path, "w") as f:
        f.write(f"{nC} {n_edges}\n")
        for c in range(nC):
            f.write(" ".join(str(int(x)) for x in coc[c][valid[c]]) + "\n")


def write_graph_part(part: np.ndarray, path: str) -> None:
    """``graph.info.part.N``: one owning part id per line, in global cell order."""
    np.savetxt(path, np.asarray(part, dtype=np.int64), fmt="%d")


def read_graph_part(path: str, n_cells: int | None = None) -> np.ndarray:
    part = np.loadtxt(path, dtype=np.int64, ndmin=1).astype(np.int32)
    if n_cells is not None and part.shape[0] != n_cells:
        raise ValueError(f"{path}: {part.shape[0]} lines, mesh has {n_cells} cells")
    return part


# ---------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------

def vertex_owner_cell(mesh) -> np.ndarray:
    """0-based global cell that owns each vertex: the first valid entry of cellsOnVertex(:, v)."""
    nC, nV = int(mesh.nCells), int(mesh.nVertices)
    cov = mesh.cellsOnVertex[:nV].astype(np.int64) - 1
    valid = (cov >= 0) & (cov < nC)
    if not np.all(valid.any(axis=1)):
        raise ValueError("a vertex without any valid cell")
    first = np.argmax(valid, axis=1)
    return cov[np.arange(nV), first]


def default_halos(mesh) -> int:
    return 1 if int(mesh.vertexDegree) == 3 else 2


def build_block(mesh, part: np.ndarray, rank: int, n_halos: int | None = None) -> Mesh:
    """The block of ``rank``: a local Mesh (same fields and conventions as the global one, junk slots
    included) plus

      nCellsSolve, nVerticesSolve        owned counts (the reference's dimension names)
      indexToCellID, indexToVertexID     1-based global ids of the local entities (MPAS names)
      cellHaloLayer                      0 for owned cells, 1.. for halo layers
      vertexOwner                        owning rank of every local vertex
    """
    nC, nV, M, D = int(mesh.nCells), int(mesh.nVertices), int(mesh.maxEdges), int(mesh.vertexDegree)
    part = np.asarray(part)
    if part.shape[0] != nC:
        raise ValueError("partition length != nCells")
    if n_halos is None:
        n_halos = default_halos(mesh)

    owned = np.nonzero(part == rank)[0].astype(np.int64)
    in_block = np.zeros(nC + 1, dtype=bool)
    in_block[owned] = True
    layers = [owned]
    frontier = owned
    slot = np.arange(M)[None, :]
    for _ in range(n_halos):
        nb = mesh.cellsOnCell[frontier].astype(np.int64) - 1
        ok = (slot < mesh.nEdgesOnCell[frontier][:, None]) & (nb >= 0) & (nb < nC)
        nb = np.unique(nb[ok])
        nb = nb[~in_block[nb]]
        in_block[nb] = True
        layers.append(nb)
        frontier = nb
    cells = np.concatenate(layers)
    nCl = cells.shape[0]
    layer_of = np.concatenate([np.full(l.shape[0], i, dtype=np.int32) for i, l in enumerate(layers)])
    g2l_cell = np.full(nC + 1, nCl, dtype=np.int64)          # anything not local -> local junk slot
    g2l_cell[cells] = np.arange(nCl)

    own_cell = vertex_owner_cell(mesh)
    own_rank = part[own_cell].astype(np.int32)
    owned_v = np.nonzero(own_rank == rank)[0].astype(np.int64)
    voc_g = mesh.verticesOnCell[cells].astype(np.int64) - 1
    vmask = slot < mesh.nEdgesOnCell[cells][:, None]
    all_v = np.unique(voc_g[vmask])
    halo_v = all_v[own_rank[all_v] != rank]
    halo_v = halo_v[np.lexsort((halo_v, own_rank[halo_v]))]  # grouped by owner, ascending id inside
    verts = np.concatenate([owned_v, halo_v])
    nVl, nVs = verts.shape[0], owned_v.shape[0]
    g2l_vert = np.full(nV + 1, nVl, dtype=np.int64)
    g2l_vert[verts] = np.arange(nVl)

    # closure of the subcycle stencil: every valid cell around an owned vertex is local
    cov_owned = mesh.cellsOnVertex[owned_v].astype(np.int64) - 1
    cv_valid = (cov_owned >= 0) & (cov_owned < nC)
    if not np.all(in_block[np.where(cv_valid, cov_owned, nC)] | ~cv_valid):
        raise ValueError(f"n_halos = {n_halos} does not close the vertex stencil (vertexDegree {D}); use more layers")

    b = Mesh()
    b.on_a_sphere = mesh.on_a_sphere
    b.sphere_radius = mesh.sphere_radius
    b.kind = mesh.get("kind")
    for k in ("Lx", "Ly", "dc", "nx", "ny", "level"):
        if k in mesh:
            b[k] = mesh[k]
    b.nCells, b.nVertices, b.maxEdges, b.vertexDegree = nCl, nVl, M, D
    b.nCellsSolve, b.nVerticesSolve = int(owned.shape[0]), int(nVs)
    b.rank, b.nRanks, b.nHalos = int(rank), int(part.max()) + 1 if nC else 1, int(n_halos)

    def cell1d(a, junk):
        out = np.empty(nCl + 1, dtype=a.dtype)
        out[:nCl] = a[cells]
        out[nCl] = junk
        return out

    def vert1d(a, junk):
        out = np.empty(nVl + 1, dtype=a.dtype)
        out[:nVl] = a[verts]
        out[nVl] = junk
        return out

    b.nEdgesOnCell = cell1d(mesh.nEdgesOnCell, 0)
    voc_l = np.full((nCl + 1, M), nVl + 1, dtype=np.int32)
    voc_l[:nCl] = np.where(vmask, g2l_vert[np.where(vmask, voc_g, nV)] + 1, nVl + 1)
    b.verticesOnCell = voc_l
    coc_g = mesh.cellsOnCell[cells].astype(np.int64) - 1
    coc_ok = vmask & (coc_g >= 0) & (coc_g < nC)
    coc_l = np.full((nCl + 1, M), nCl + 1, dtype=np.int32)
    coc_l[:nCl] = np.where(coc_ok, g2l_cell[np.where(coc_ok, coc_g, nC)] + 1, nCl + 1)
    b.cellsOnCell = coc_l
    cov_g = mesh.cellsOnVertex[verts].astype(np.int64) - 1
    cov_ok = (cov_g >= 0) & (cov_g < nC)
    cov_l = np.full((nVl + 1, D), nCl + 1, dtype=np.int32)
    cov_l[:nVl] = np.where(cov_ok, g2l_cell[np.where(cov_ok, cov_g, nC)] + 1, nCl + 1)
    b.cellsOnVertex = cov_l

    # edges of the local cells (dvEdge is read through edgesOnCell by the PWL basis, pwl.F:161-176)
    if "edgesOnCell" in mesh:
        eoc_g = mesh.edgesOnCell[cells].astype(np.int64) - 1
        edges = np.unique(eoc_g[vmask])
        nEl = edges.shape[0]
        g2l_edge = np.full(int(mesh.nEdges) + 1, nEl, dtype=np.int64)
        g2l_edge[edges] = np.arange(nEl)
        eoc_l = np.full((nCl + 1, M), nEl + 1, dtype=np.int32)
        eoc_l[:nCl] = np.where(vmask, g2l_edge[np.where(vmask, eoc_g, int(mesh.nEdges))] + 1, nEl + 1)
        b.edgesOnCell = eoc_l
        b.nEdges = nEl
        for k in ("dvEdge", "dcEdge"):
            if k in mesh:
                out = np.zeros(nEl + 1)
                out[:nEl] = mesh[k][edges]
                b[k] = out
        b.indexToEdgeID = (edges + 1).astype(np.int32)
        if "cellsOnEdge" in mesh:          # weak stress divergence reads the two cells of an edge (weak.F:585-592)
            coe_g = mesh.cellsOnEdge[edges].astype(np.int64) - 1
            coe_ok = (coe_g >= 0) & (coe_g < nC)
            coe_l = np.full((nEl + 1, 2), nCl + 1, dtype=np.int32)
            coe_l[:nEl] = np.where(coe_ok, g2l_cell[np.where(coe_ok, coe_g, nC)] + 1, nCl + 1)
            b.cellsOnEdge = coe_l

    for k in CELL_FIELDS_1D:
        b[k] = cell1d(mesh[k], JUNK_AREA if k == "areaCell" else 0.0)
    for k in VERTEX_FIELDS_1D:
        b[k] = vert1d(mesh[k], 0.0)
    kite = np.zeros((nVl + 1, D))
    kite[:nVl] = mesh.kiteAreasOnVertex[verts]
    b.kiteAreasOnVertex = kite

    b.indexToCellID = (cells + 1).astype(np.int32)
    b.indexToVertexID = (verts + 1).astype(np.int32)
    b.cellHaloLayer = layer_of
    b.cellOwner = part[cells].astype(np.int32)              # owning rank of every local cell
    b.vertexOwner = np.concatenate([np.full(nVs, rank, dtype=np.int32), own_rank[halo_v]])
    # interiorVertex is computed on owned vertices and halo-exchanged by the reference (mesh.F:405);
    # taking it from the global mesh is that exchange
    cov_all = mesh.cellsOnVertex[verts]
    interior = np.zeros(nVl + 1, dtype=np.int32)
    interior[:nVl] = np.all((cov_all >= 1) & (cov_all <= nC), axis=1)
    b.interiorVertex = interior
    return b


# ---------------------------------------------------------------------------------------------
# exchange lists
# ---------------------------------------------------------------------------------------------

def halo_requests(block) -> dict:
    """{owner rank: global 1-based vertex ids this block needs from it}, in local halo order (which is
    ascending global id inside each owner group)."""
    nVs, nVl = int(block.nVerticesSolve), int(block.nVertices)
    owner = block.vertexOwner[nVs:nVl]
    gid = block.indexToVertexID[nVs:nVl]
    out = {}
    for r in np.unique(owner):
        out[int(r)] = gid[owner == r].copy()
    return out


def exchange_lists(block, requests_by_rank: dict):
    """The evp_set_halo() arguments of ``block``.

    requests_by_rank[q] = halo_requests(block of rank q) for every rank q (at least for the ranks
    that need something from this one).  Returns (neighbourRank, sendOffset, sendIndex, recvOffset,
    recvIndex), indices 1-based local, neighbours ascending; the k-th value sent to a neighbour is the
    k-th value that neighbour expects (both sides order by global id)."""
    rank = int(block.rank)
    nVs, nVl = int(block.nVerticesSolve), int(block.nVertices)
    mine = requests_by_rank.get(rank)
    if mine is None:
        mine = halo_requests(block)
    gid_owned = block.indexToVertexID[:nVs].astype(np.int64)        # ascending by construction
    send = {}
    for q, req in requests_by_rank.items():
        if int(q) == rank or req is None:
            continue
        want = req.get(rank)
        if want is None or len(want) == 0:
            continue
        want = np.asarray(want, dtype=np.int64)
        pos = np.searchsorted(gid_owned, want)
        if np.any(pos >= nVs) or np.any(gid_owned[np.minimum(pos, nVs - 1)] != want):
            raise ValueError(f"rank {q} requests vertices rank {rank} does not own")
        send[int(q)] = (pos + 1).astype(np.int32)
    nbrs = sorted(set(send) | set(int(r) for r in mine))
    send_off, recv_off = [0], [0]
    send_idx, recv_idx = [], []
    gid_halo = block.indexToVertexID[nVs:nVl]
    owner = block.vertexOwner[nVs:nVl]
    for q in nbrs:
        s = send.get(q, np.zeros(0, dtype=np.int32))
        r = (np.nonzero(owner == q)[0] + nVs + 1).astype(np.int32)
        if q in mine:
            assert np.array_equal(gid_halo[r - nVs - 1], np.asarray(mine[q]))
        send_idx.append(s)
        recv_idx.append(r)
        send_off.append(send_off[-1] + s.shape[0])
        recv_off.append(recv_off[-1] + r.shape[0])
    cat = lambda l: np.concatenate(l).astype(np.int32) if l else np.zeros(0, dtype=np.int32)
    return (np.asarray(nbrs, dtype=np.int32), np.asarray(send_off, dtype=np.int32), cat(send_idx),
            np.asarray(recv_off, dtype=np.int32), cat(recv_idx))


def cell_halo_requests(block) -> dict:
    """{owner rank: ascending global ids (1-based) of the halo CELLS this block needs from it} -- the cell
    counterpart of halo_requests, for fields that live on cells (the tracers of the transport scheme)."""
    nCs, nCl = int(block.nCellsSolve), int(block.nCells)
    gid, owner = block.indexToCellID[nCs:nCl].astype(np.int64), block.cellOwner[nCs:nCl]
    return {int(q): np.sort(gid[owner == q]) for q in np.unique(owner)}


def cell_exchange_lists(block, requests_by_rank: dict):
    """(neighbours, send_off, send_idx, recv_off, recv_idx) with 1-based LOCAL cell indices: what this rank packs
    for every neighbour (its owned cells, in the ascending-global-id order the neighbour asked for) and where the
    values it receives go (its halo cells owned by that neighbour, same order)."""
    rank = int(block.rank)
    nCs, nCl = int(block.nCellsSolve), int(block.nCells)
    gid = block.indexToCellID.astype(np.int64)
    order_owned = np.argsort(gid[:nCs], kind="stable")
    gid_owned = gid[:nCs][order_owned]
    mine = cell_halo_requests(block)
    send = {}
    for q, req in requests_by_rank.items():
        if int(q) == rank or req is None:
            continue
        want = req.get(rank)
        if want is None or len(want) == 0:
            continue
        want = np.asarray(want, dtype=np.int64)
        pos = np.searchsorted(gid_owned, want)
        if np.any(pos >= nCs) or np.any(gid_owned[np.minimum(pos, nCs - 1)] != want):
            raise ValueError(f"rank {q} requests cells rank {rank} does not own")
        send[int(q)] = (order_owned[pos] + 1).astype(np.int32)
    nbrs = sorted(set(send) | set(mine))
    send_off, recv_off, send_idx, recv_idx = [0], [0], [], []
    halo_gid, halo_owner = gid[nCs:nCl], block.cellOwner[nCs:nCl]
    for q in nbrs:
        s = send.get(q, np.zeros(0, dtype=np.int32))
        loc = np.nonzero(halo_owner == q)[0]
        loc = loc[np.argsort(halo_gid[loc], kind="stable")]
        r = (loc + nCs + 1).astype(np.int32)
        send_idx.append(s)
        recv_idx.append(r)
        send_off.append(send_off[-1] + s.shape[0])
        recv_off.append(recv_off[-1] + r.shape[0])
    cat = lambda l: np.concatenate(l).astype(np.int32) if l else np.zeros(0, dtype=np.int32)
    return (np.asarray(nbrs, dtype=np.int32), np.asarray(send_off, dtype=np.int32), cat(send_idx),
            np.asarray(recv_off, dtype=np.int32), cat(recv_idx))


# ---------------------------------------------------------------------------------------------
# moving fields between the global mesh and a block
# ---------------------------------------------------------------------------------------------

def restrict_field(block, a: np.ndarray, n_global_cells: int, n_global_vertices: int) -> np.ndarray:
    """Global field (with junk slot) -> local field (with junk slot = 0)."""
    if a.shape[0] == n_global_cells + 1:
        idx = block.indexToCellID.astype(np.int64) - 1
    elif a.shape[0] == n_global_vertices + 1:
        idx = block.indexToVertexID.astype(np.int64) - 1
    else:
        raise ValueError(f"field of length {a.shape[0]} is neither a cell nor a vertex field")
    out = np.zeros((idx.shape[0] + 1,) + a.shape[1:], dtype=a.dtype)
    out[:-1] = a[idx]
    return out


def restrict_weak(block, gmesh, gweak: dict) -> dict:
    """The weak-operator fields of weakmesh.weak_fields(global mesh) restricted to ``block``: cell- and
    vertex-indexed arrays by rows, edge-indexed ones through the block's edge numbering.  Restricting the global
    arrays (instead of recomputing them per block) keeps every value bit-identical on every rank count."""
    nCg, nVg, nEg = int(gmesh.nCells), int(gmesh.nVertices), int(gmesh.nEdges)
    nVl, nEl = int(block.nVertices), int(block.nEdges)
    out = {}
    for k in ("normalVectorPolygon", "normalVectorTriangle", "latCellRotated", "latVertexRotated"):
        out[k] = np.ascontiguousarray(restrict_field(block, gweak[k], nCg, nVg))
    g2l_vert = np.full(nVg + 2, nVl + 1, dtype=np.int64)
    g2l_vert[block.indexToVertexID.astype(np.int64)] = np.arange(1, nVl + 1)
    g2l_edge = np.full(nEg + 2, nEl + 1, dtype=np.int64)
    g2l_edge[block.indexToEdgeID.astype(np.int64)] = np.arange(1, nEl + 1)
    edges_g = block.indexToEdgeID.astype(np.int64) - 1
    voe = np.full((nEl + 1, 2), nVl + 1, dtype=np.int32)
    voe[:nEl] = g2l_vert[gweak["verticesOnEdge"][edges_g].astype(np.int64)].astype(np.int32)
    out["verticesOnEdge"] = voe
    verts_g = block.indexToVertexID.astype(np.int64) - 1
    eov = np.full((nVl + 1, gweak["edgesOnVertex"].shape[1]), nEl + 1, dtype=np.int32)
    eov[:nVl] = g2l_edge[gweak["edgesOnVertex"][verts_g].astype(np.int64)].astype(np.int32)
    out["edgesOnVertex"] = eov
    return out


def restrict_ir(block, gmesh, girf: dict) -> dict:
    """irmesh.ir_fields(global mesh) restricted to ``block``: the mesh-file arrays (verticesOnEdge, edgesOnVertex,
    x/y/zEdge) through the block's numbering, and coeffs_reconstruct by rows -- the reference computes it on owned
    cells and halo-exchanges it (incremental_remap.F:744-770), which restricting the global array reproduces."""
    nVl, nEl = int(block.nVertices), int(block.nEdges)
    nVg, nEg = int(gmesh.nVertices), int(gmesh.nEdges)
    verts = block.indexToVertexID.astype(np.int64) - 1
    edges = block.indexToEdgeID.astype(np.int64) - 1
    g2l_vert = np.full(nVg + 1, nVl, dtype=np.int64)
    g2l_vert[verts] = np.arange(nVl)
    g2l_edge = np.full(nEg + 1, nEl, dtype=np.int64)
    g2l_edge[edges] = np.arange(nEl)
    voe = np.full((nEl + 1, 2), nVl + 1, dtype=np.int32)
    voe[:nEl] = g2l_vert[girf["verticesOnEdge"][edges].astype(np.int64) - 1] + 1
    eov = np.full((nVl + 1, girf["edgesOnVertex"].shape[1]), nEl + 1, dtype=np.int32)
    eov[:nVl] = g2l_edge[np.minimum(girf["edgesOnVertex"][verts].astype(np.int64) - 1, nEg)] + 1
    out = dict(verticesOnEdge=voe, edgesOnVertex=eov)
    for k in ("xEdge", "yEdge", "zEdge"):
        a = np.zeros(nEl + 1)
        a[:nEl] = girf[k][edges]
        out[k] = a
    out["coeffs_reconstruct"] = restrict_field(block, girf["coeffs_reconstruct"], int(gmesh.nCells), nVg)
    return out


def restrict_step(block, step: dict, n_global_cells: int, n_global_vertices: int) -> dict:
    """Per-step fields of the global mesh -> the block (what the reference's halo exchanges of the
    pre-subcycle leave in owned + halo entries)."""
    out = {}
    for k, v in step.items():
        out[k] = restrict_field(block, v, n_global_cells, n_global_vertices) if isinstance(v, np.ndarray) else v
    return out


def scatter_owned(block, local: np.ndarray, global_out: np.ndarray, kind: str) -> None:
    """Write the OWNED entries of a local field into the global field (gathering a decomposed result)."""
    if kind == "cell":
        n, idx = int(block.nCellsSolve), block.indexToCellID
    else:
        n, idx = int(block.nVerticesSolve), block.indexToVertexID
    global_out[idx[:n].astype(np.int64) - 1] = local[:n]
