"""Decomposition, halo maps and exchange lists (host logic, CPU only).

The reference pins none of these maps numerically (they live in the un-vendored MPAS framework and in
METIS output): SURVEY.md 8(c) -> "parity unpinned" for the maps themselves.  What IS asserted:
self-consistency of the maps, and the reference's own parallelism policy
(testing_and_setup/testing/tests/parallelism.py:75-85): owned results are BIT-identical for every
rank count."""
import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import partition


@pytest.mark.parametrize("kind,n_parts,method", [("hex20", 2, "rcb"), ("hex20", 5, "rcb"), ("ico3", 4, "block"),
                                                 ("ico3", 8, "rcb"), ("quad40", 3, "rcb"), ("ico3", 1, "block")])
def test_blocks_are_consistent(kind, n_parts, method):
    mesh, _ = common.mesh_case(kind)
    nC, nV, M, D = mesh.nCells, mesh.nVertices, mesh.maxEdges, mesh.vertexDegree
    part, blocks, lists = common.make_blocks(mesh, n_parts, method)
    counts = np.bincount(part, minlength=n_parts)
    assert counts.max() - counts.min() <= 1                      # balanced
    assert np.array_equal(part, partition.partition_cells(mesh, n_parts, method))   # deterministic
    cell_owned = np.zeros(nC, dtype=int)
    vert_owned = np.zeros(nV, dtype=int)
    for b in blocks:
        nCs, nVs = b.nCellsSolve, b.nVerticesSolve
        gc = b.indexToCellID.astype(np.int64) - 1
        gv = b.indexToVertexID.astype(np.int64) - 1
        cell_owned[gc[:nCs]] += 1
        vert_owned[gv[:nVs]] += 1
        assert np.all(np.diff(gc[:nCs]) > 0) and np.all(np.diff(gv[:nVs]) > 0)
        assert len(np.unique(gc)) == len(gc) and len(np.unique(gv)) == len(gv)
        assert np.all(part[gc[:nCs]] == b.rank) and np.all(part[gc[nCs:]] != b.rank)
        # connectivity maps back to the global one, slot by slot
        n = b.nEdgesOnCell[:b.nCells]
        assert np.array_equal(n, mesh.nEdgesOnCell[gc])
        assert b.nEdgesOnCell[b.nCells] == 0
        for s in range(M):
            ok = n > s
            assert np.array_equal(gv[b.verticesOnCell[:b.nCells][ok, s] - 1], mesh.verticesOnCell[gc[ok], s] - 1)
        cov_l = b.cellsOnVertex[:b.nVertices]
        cov_g = mesh.cellsOnVertex[gv]
        local = cov_l <= b.nCells
        assert np.array_equal(gc[cov_l[local] - 1], cov_g[local] - 1)
        # a cell that is not local is either the global junk cell or a non-local cell
        assert np.all(cov_l[~local] == b.nCells + 1)
        # stencil closure for owned vertices
        own_valid = (cov_g[:nVs] >= 1) & (cov_g[:nVs] <= nC)
        assert np.all(local[:nVs] == own_valid)
        # geometry travels
        assert np.array_equal(b.xVertex[:b.nVertices], mesh.xVertex[gv])
        assert np.array_equal(b.areaCell[:b.nCells], mesh.areaCell[gc]) and b.areaCell[b.nCells] == mesh.areaCell[nC]
    assert np.all(cell_owned == 1) and np.all(vert_owned == 1)
    # exchange lists: what q sends to r is exactly what r expects from q, in the same order
    for r, (nbr, soff, sidx, roff, ridx) in enumerate(lists):
        b = blocks[r]
        assert np.all(np.diff(nbr) > 0) and r not in nbr
        assert np.all(sidx >= 1) and np.all(sidx <= b.nVerticesSolve)
        assert np.all(ridx > b.nVerticesSolve) and np.all(ridx <= b.nVertices)
        # every halo vertex is received exactly once; the recv slices are contiguous and in order
        assert np.array_equal(ridx, np.arange(b.nVerticesSolve + 1, b.nVertices + 1))
        for k, q in enumerate(nbr):
            qn, qsoff, qsidx, qroff, qridx = lists[q]
            kk = int(np.nonzero(qn == r)[0][0])
            sent = blocks[q].indexToVertexID[qsidx[qsoff[kk]:qsoff[kk + 1]] - 1]
            want = b.indexToVertexID[ridx[roff[k]:roff[k + 1]] - 1]
            assert np.array_equal(sent, want)


def test_too_few_halos_is_rejected():
    mesh, _ = common.mesh_case("quad40")
    part = partition.partition_cells(mesh, 4, "rcb")
    with pytest.raises(ValueError, match="halo"):
        partition.build_block(mesh, part, 0, n_halos=1)


def test_graph_files_roundtrip(tmp_path):
    mesh, _ = common.mesh_case("hex20")
    part = partition.partition_cells(mesh, 3, "rcb")
    partition.write_graph_info(mesh, str(tmp_path / "graph.info"))
    partition.write_graph_part(part, str(tmp_path / "graph.info.part.3"))
    assert np.array_equal(partition.read_graph_part(str(tmp_path / "graph.info.part.3"), mesh.nCells), part)
    with open(tmp_path / "graph.info") as f:
        head = f.readline().split()
        lines = f.read().strip().split("\n")
    assert int(head[0]) == mesh.nCells == len(lines)
    deg = sum(len(l.split()) for l in lines)
    assert deg == 2 * int(head[1])
    with pytest.raises(ValueError):
        partition.read_graph_part(str(tmp_path / "graph.info.part.3"), mesh.nCells + 1)


@pytest.mark.parametrize("kind,n_parts,method,nsub", [("hex20", 2, "rcb", 12), ("hex20", 3, "block", 12),
                                                      ("ico3", 4, "block", 12), ("ico3", 3, "rcb", 12),
                                                      ("quad40", 2, "rcb", 6)])
def test_owned_results_bit_identical_across_rank_counts(kind, n_parts, method, nsub):
    """The reference's parallelism policy with the oracle as the per-block solver."""
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    out, blocks, _ = common.run_oracle_blocks(mesh, step, opts, nsub, n_parts, method)
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_CELL:
        assert np.array_equal(out[k][cm], ref[k][cm]), k
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(out[k][vm], ref[k][vm]), k
    assert np.abs(ref["uVelocity"]).max() > 0


@pytest.mark.parametrize("kind,n_parts,method,schemes", [("ico3", 3, "rcb", ("weak", "weak")),
                                                         ("hex20", 2, "block", ("weak", "weak")),
                                                         ("quad40", 2, "rcb", ("weak", "weak"))])
def test_weak_operators_bit_identical_across_rank_counts(kind, n_parts, method, schemes):
    """The weak operators decomposed: edge-indexed mesh data goes through the block's edge numbering
    (partition.build_block: edgesOnCell, cellsOnEdge, dvEdge, dcEdge; partition.restrict_weak: verticesOnEdge,
    edgesOnVertex, the normal vectors); owned results equal the single-rank run bit for bit."""
    from mpas_seaice_b200 import partition, weakmesh
    mesh, var = common.mesh_case(kind)
    gweak = weakmesh.weak_fields(mesh)
    step, opts = common.step_case(mesh)
    opts = dict(opts, strain_scheme=schemes[0], stress_divergence_scheme=schemes[1])
    nsub = 8
    ref = common.run_oracle(mesh, dict(var, weak=gweak), step, opts, nsub)
    part, blocks, lists = common.make_blocks(mesh, n_parts, method)
    bsteps = [partition.restrict_step(b, step, mesh.nCells, mesh.nVertices) for b in blocks]
    bvars = []
    for b in blocks:
        v = oracle.init_variational(b)
        v["weak"] = partition.restrict_weak(b, mesh, gweak)
        bvars.append(v)
    for _ in range(nsub):
        for b, v, s in zip(blocks, bvars, bsteps):
            oracle.subcycle_velocity_solver(b, v, s, dict(opts, nVerticesSolve=int(b.nVerticesSolve)), 1)
        common.exchange_halos(bsteps, lists)
    for b, s in zip(blocks, bsteps):
        nVs, nCs = int(b.nVerticesSolve), int(b.nCellsSolve)
        gv = b.indexToVertexID[:nVs].astype(np.int64) - 1
        gc = b.indexToCellID[:nCs].astype(np.int64) - 1
        for k in ("uVelocity", "vVelocity", "stressDivergenceU", "stressDivergenceV"):
            assert np.array_equal(s[k][:nVs], ref[k][gv]), k
        for k in ("stress11Weak", "stress22Weak", "stress12Weak", "strain11Weak", "strain12Weak"):
            assert np.array_equal(s[k][:nCs], ref[k][gc]), k
    assert np.abs(ref["uVelocity"]).max() > 0


@pytest.mark.parametrize("kind,n_parts,n_halos", [("ico4", 4, 2), ("hex20", 3, 2), ("quad40", 3, 3)])
def test_cell_exchange_lists_are_consistent(kind, n_parts, n_halos):
    """The cell halo maps of the transport (partition.cell_halo_requests / cell_exchange_lists): every halo cell is
    received exactly once, from the rank that owns it; what a rank sends are owned cells; sender and receiver agree on
    the order (ascending global id).  The reference's own maps live in the MPAS framework (unpinned): self-consistency
    plus identical results across rank counts (tests/test_ir_multirank.py) is the check."""
    mesh, _ = common.mesh_case(kind)
    part = partition.partition_cells(mesh, n_parts)
    blocks = [partition.build_block(mesh, part, r, n_halos) for r in range(n_parts)]
    requests = {r: partition.cell_halo_requests(b) for r, b in enumerate(blocks)}
    lists = [partition.cell_exchange_lists(b, requests) for b in blocks]
    for r, b in enumerate(blocks):
        nbrs, soff, sidx, roff, ridx = lists[r]
        assert np.array_equal(b.cellOwner[:b.nCellsSolve], np.full(b.nCellsSolve, r))
        got = np.sort(ridx)
        assert np.array_equal(got, np.arange(b.nCellsSolve + 1, b.nCells + 1))          # every halo cell, once
        assert np.all(sidx >= 1) and np.all(sidx <= b.nCellsSolve)                        # only owned cells are sent
        for k, q in enumerate(nbrs):
            q = int(q)
            qn, qsoff, qsidx, qroff, qridx = lists[q]
            kk = int(np.nonzero(qn == r)[0][0])
            mine = ridx[roff[k]:roff[k + 1]] - 1
            theirs = qsidx[qsoff[kk]:qsoff[kk + 1]] - 1
            assert np.array_equal(b.indexToCellID[mine], blocks[q].indexToCellID[theirs])
            assert np.all(b.cellOwner[mine] == q)


@pytest.mark.parametrize("seed", range(8))
def test_random_cell_assignments_reproduce_the_single_block_run(seed):
    """Any cell-to-rank assignment must do -- the reference takes whatever graph.info.part.N holds: every cell drawn at
    random, a partition with a tenth of its cells reassigned, a part of two cells, index runs of random lengths (some
    empty).  Blocks, halo lists and the decomposed run against the single-block oracle, bit for bit.  (144 seeds when this
    was written: no difference.)"""
    import oracle
    rng = np.random.default_rng(31000 + seed)
    mesh, var = common.mesh_case(["hex20", "ico3", "quad40"][seed % 3])
    nC, nV = mesh.nCells, mesh.nVertices
    n_parts = int(rng.integers(2, 7))
    mode = seed % 4
    if mode == 0:
        part = rng.integers(0, n_parts, nC)
    elif mode == 1:
        part = partition.partition_cells(mesh, n_parts, "rcb").copy()
        flip = rng.uniform(size=nC) < 0.1
        part[flip] = rng.integers(0, n_parts, int(flip.sum()))
    elif mode == 2:
        part = partition.partition_cells(mesh, n_parts - 1, "block").copy() if n_parts > 2 else np.zeros(nC, int)
        part[rng.integers(0, nC, 2)] = n_parts - 1
    else:
        part = np.searchsorted(np.sort(rng.integers(0, nC, n_parts - 1)), np.arange(nC), side="right")
    part = np.asarray(part, dtype=np.int64)
    step, opts = common.step_case(mesh)
    step["solveStress"][:nC][rng.uniform(size=nC) < 0.2] = 0
    step["solveVelocity"][:nV][rng.uniform(size=nV) < 0.2] = 0
    n_sub = int(rng.integers(2, 6))
    ref = common.run_oracle(mesh, var, step, opts, n_sub)
    blocks = [partition.build_block(mesh, part, r, None) for r in range(n_parts)]
    requests = {r: partition.halo_requests(b) for r, b in enumerate(blocks)}
    lists = [partition.exchange_lists(b, requests) for b in blocks]
    bvars = [oracle.init_variational(b) if b.nCells > 0 else None for b in blocks]
    bsteps = [partition.restrict_step(b, step, nC, nV) for b in blocks]
    for _ in range(n_sub):
        for b, v, s in zip(blocks, bvars, bsteps):
            if b.nCells > 0:
                oracle.subcycle_velocity_solver(b, v, s, dict(opts, nVerticesSolve=int(b.nVerticesSolve)), 1)
        common.exchange_halos(bsteps, lists)
    out = {k: np.zeros_like(step[k]) for k in common.COMPARE_CELL + common.COMPARE_VERTEX}
    for b, s in zip(blocks, bsteps):
        for k in common.COMPARE_CELL:
            partition.scatter_owned(b, s[k], out[k], "cell")
        for k in common.COMPARE_VERTEX:
            partition.scatter_owned(b, s[k], out[k], "vertex")
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_CELL:
        assert np.array_equal(out[k][cm], ref[k][cm]), k
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(out[k][vm], ref[k][vm]), k


def test_weak_strain_variational_divergence_is_rank_count_dependent_in_the_reference():
    """config_strain_scheme = 'weak' with config_stress_divergence_scheme = 'variational' is NOT decomposition-invariant in
    the reference: interpolate_strains_weak_to_variational fills strain11/22/12Vertex for iVertex <= nVerticesSolve only
    (velocity_solver.F:2933) and nothing updates their halo (Registry.xml:3769-3771 are never exchanged), so an owned
    cell reads a stale vertex strain at every vertex another rank owns.  The oracle restates the routine as written, block
    by block; this test pins that behaviour (a decomposed run of this mix differs from the single-block run, starting at
    the block boundary) so that nobody 'fixes' it on one side only.  The weak / weak pair above and every
    variational configuration are invariant."""
    from mpas_seaice_b200 import weakmesh
    mesh, var = common.mesh_case("hex20")
    gweak = weakmesh.weak_fields(mesh)
    step, opts = common.step_case(mesh)
    opts = dict(opts, strain_scheme="weak", stress_divergence_scheme="variational")
    nsub = 3
    ref = common.run_oracle(mesh, dict(var, weak=gweak), step, opts, nsub)
    part, blocks, lists = common.make_blocks(mesh, 2, "block", n_halos=2)
    bsteps = [partition.restrict_step(b, step, mesh.nCells, mesh.nVertices) for b in blocks]
    bvars = []
    for b in blocks:
        v = oracle.init_variational(b)
        v["weak"] = partition.restrict_weak(b, mesh, gweak)
        bvars.append(v)
    for _ in range(nsub):
        for b, v, s in zip(blocks, bvars, bsteps):
            oracle.subcycle_velocity_solver(b, v, s, dict(opts, nVerticesSolve=int(b.nVerticesSolve)), 1)
        common.exchange_halos(bsteps, lists)
    differing, total = 0, 0
    for r, (b, s) in enumerate(zip(blocks, bsteps)):
        nCs = int(b.nCellsSolve)
        gc = b.indexToCellID[:nCs].astype(np.int64) - 1
        bad = np.any(s["stress11"][:nCs] != ref["stress11"][gc], axis=1)
        differing += int(bad.sum())
        total += nCs
    assert 0 < differing < total            # next to the block boundary, spreading a ring per subcycle; the far side is untouched
