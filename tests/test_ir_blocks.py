"""Incremental remapping on a decomposed mesh: blocks with two halo layers (what check_halo_layer_number asks for on
hexagonal meshes, incremental_remap.F:829-870), transport on the owned cells of every block, tracer halo update
(seaice_update_tracer_halo, :2705-2712) -- bit-identical to the single-block run, which is the reference's
regression policy across rank counts (testing_and_setup/testing: parallelism test).  Runs the oracle and, through
the C ABI, the kernels under host emulation (tests/test_ir_parity.py explains the two legs)."""
import numpy as np
import pytest

from oracle import ir
from mpas_seaice_b200 import ir_host, partition
from test_oracle_ir import case, smooth_divergent_velocity, _random_state
from test_ir_parity import clone, lib_path  # noqa: F401  (lib_path: the emulation / cuda fixture)


def _blocks(kind, n_parts, n_halos):
    mesh, irf, geom = case(kind)
    part = partition.partition_cells(mesh, n_parts)
    blocks = [partition.build_block(mesh, part, r, n_halos) for r in range(n_parts)]
    birfs = [partition.restrict_ir(b, mesh, irf) for b in blocks]
    return mesh, irf, geom, blocks, birfs


def _halo_update(mesh, blocks, block_tracers, global_tracers):
    """owned values -> global arrays -> every block's copy (owner to halo)"""
    for t in range(len(global_tracers)):
        for b, tr in zip(blocks, block_tracers):
            partition.scatter_owned(b, tr[t].array, global_tracers[t].array, "cell")
        for b, tr in zip(blocks, block_tracers):
            tr[t].array[:] = partition.restrict_field(b, global_tracers[t].array, mesh.nCells, mesh.nVertices)


@pytest.mark.parametrize("kind,n_parts", [("ico4", 3), ("hex16", 2)])
def test_oracle_blocks_reproduce_the_single_block_run(kind, n_parts):
    mesh, irf, geom, blocks, birfs = _blocks(kind, n_parts, 2)
    rng = np.random.default_rng(17)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    single = clone(tracers)
    gathered = clone(tracers)
    btr = [[ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like)
            for t in tracers] for b in blocks]
    bgeom = [ir.init_geometry(b, f, n_cells_solve=b.nCellsSolve) for b, f in zip(blocks, birfs)]
    bu = [partition.restrict_field(b, u, mesh.nCells, mesh.nVertices) for b in blocks]
    bv = [partition.restrict_field(b, v, mesh.nCells, mesh.nVertices) for b in blocks]
    for _ in range(3):
        ir.run(mesh, irf, geom, single, u, v, 3600.0)
        for b, f, g, tr, uu, vv in zip(blocks, birfs, bgeom, btr, bu, bv):
            ir.run(b, f, g, tr, uu, vv, 3600.0, n_cells_solve=b.nCellsSolve)
        _halo_update(mesh, blocks, btr, gathered)
    for a, g in zip(single, gathered):
        assert np.array_equal(a.array[:mesh.nCells], g.array[:mesh.nCells]), a.name


def test_one_halo_layer_is_not_enough():
    """With a single halo layer the side cells' gradients miss a neighbour: the result differs -- why the reference
    warns about config_num_halos (:845-870)."""
    mesh, irf, geom, blocks, birfs = _blocks("ico4", 3, 1)
    rng = np.random.default_rng(17)
    tracers = _random_state(mesh, rng, n_cat=1, n_ice=1, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    single, gathered = clone(tracers), clone(tracers)
    ir.run(mesh, irf, geom, single, u, v, 3600.0)
    for b, f in zip(blocks, birfs):
        tr = [ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like)
              for t in tracers]
        g = ir.init_geometry(b, f, n_cells_solve=b.nCellsSolve, check=False)
        ir.run(b, f, g, tr, partition.restrict_field(b, u, mesh.nCells, mesh.nVertices),
               partition.restrict_field(b, v, mesh.nCells, mesh.nVertices), 3600.0, n_cells_solve=b.nCellsSolve, check=False)
        for t in range(len(tr)):
            partition.scatter_owned(b, tr[t].array, gathered[t].array, "cell")
    assert not np.array_equal(single[0].array, gathered[0].array)
    assert np.allclose(single[0].array, gathered[0].array, atol=5e-2)


def test_device_blocks_reproduce_the_single_block_run(lib_path):
    """The same through ir_init_geometry / ir_run with nCellsSolve < nCells: halo cells keep their values, owned
    cells match the single-block oracle run bit for bit."""
    mesh, irf, geom, blocks, birfs = _blocks("ico4", 3, 2)
    rng = np.random.default_rng(17)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    single, gathered = clone(tracers), clone(tracers)
    solvers, btr, buv = [], [], []
    try:
        for b, f in zip(blocks, birfs):
            g = ir_host.init_geometry(b, f, n_cells_solve=b.nCellsSolve, lib_path=lib_path)
            s = ir_host.IrTransport(b, f, g, 2, n_cells_solve=b.nCellsSolve, lib_path=lib_path)
            tr = [ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like)
                  for t in tracers]
            s.set_tracers(tr)
            solvers.append(s)
            btr.append(tr)
            buv.append((partition.restrict_field(b, u, mesh.nCells, mesh.nVertices),
                        partition.restrict_field(b, v, mesh.nCells, mesh.nVertices)))
        for _ in range(2):
            ir.run(mesh, irf, geom, single, u, v, 3600.0)
            for b, s, tr, (uu, vv) in zip(blocks, solvers, btr, buv):
                halo_before = [t.array[b.nCellsSolve:b.nCells].copy() for t in tr]
                s.run(tr, uu, vv, 3600.0)
                for t, before in zip(tr, halo_before):
                    if not t.volume_like:                 # (volumes go through v / a * a on halo cells too)
                        assert np.array_equal(t.array[b.nCellsSolve:b.nCells], before), t.name
            _halo_update(mesh, blocks, btr, gathered)
    finally:
        for s in solvers:
            s.destroy()
    for a, g in zip(single, gathered):
        assert np.array_equal(a.array[:mesh.nCells], g.array[:mesh.nCells]), a.name


@pytest.mark.parametrize("seed", [0, 1, 2, 5, 8])
def test_random_cell_assignments_reproduce_the_single_block_run(seed):
    """Blocks of ANY cell-to-rank assignment (a partition with scattered cells reassigned, index runs, every cell at
    random), with the halo layers check_halo_layer_number asks for: two on vertexDegree-3 meshes, THREE on quadrilaterals
    (incremental_remap.F:848-852 -- a diagonal source cell's gradient reaches three edge steps from the owned cell; with
    two layers the scattered assignments of seeds 5 and 8 differ from the single-block run, as they should).  Oracle blocks
    + tracer halo update against the single block, bit for bit; 60 seeds when this was written."""
    rng = np.random.default_rng(41000 + seed)
    kind = ["hex16", "ico3", "quad16"][seed % 3]
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    n_parts = int(rng.integers(2, 5))
    mode = seed % 3
    if mode == 0:
        part = partition.partition_cells(mesh, n_parts).copy()
        flip = rng.uniform(size=nC) < 0.08
        part[flip] = rng.integers(0, n_parts, int(flip.sum()))
    elif mode == 1:
        part = np.searchsorted(np.sort(rng.integers(1, nC - 1, n_parts - 1)), np.arange(nC), side="right")
    else:
        part = rng.integers(0, n_parts, nC)
    part = np.asarray(part, dtype=np.int64)
    tracers = _random_state(mesh, rng, n_cat=int(rng.integers(1, 3)), n_ice=int(rng.integers(1, 3)), n_snow=int(rng.integers(0, 2)))
    u, v = smooth_divergent_velocity(mesh, geom, cfl=rng.uniform(0.1, 0.5))

    def decomposed(n_halos):
        blocks = [partition.build_block(mesh, part, r, n_halos) for r in range(n_parts)]
        birfs = [partition.restrict_ir(b, mesh, irf) for b in blocks]
        gathered = clone(tracers)
        btr = [[ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like)
                for t in tracers] for b in blocks]
        live = [b.nCellsSolve > 0 for b in blocks]
        bgeom = [ir.init_geometry(b, f, n_cells_solve=b.nCellsSolve, check=False) if ok else None
                 for b, f, ok in zip(blocks, birfs, live)]
        for _ in range(2):
            for b, f, g, tr, ok in zip(blocks, birfs, bgeom, btr, live):
                if ok:
                    ir.run(b, f, g, tr, partition.restrict_field(b, u, mesh.nCells, mesh.nVertices),
                           partition.restrict_field(b, v, mesh.nCells, mesh.nVertices), 3600.0, n_cells_solve=b.nCellsSolve,
                           check=False)
            _halo_update(mesh, blocks, btr, gathered)
        return gathered

    single = clone(tracers)
    for _ in range(2):
        ir.run(mesh, irf, geom, single, u, v, 3600.0)
    quads = int(mesh.vertexDegree) == 4
    for a, g in zip(single, decomposed(3 if quads else 2)):
        assert np.array_equal(a.array[:nC], g.array[:nC]), a.name
    if quads and seed in (5, 8):
        two = decomposed(2)
        assert not all(np.array_equal(a.array[:nC], g.array[:nC]) for a, g in zip(single, two))
