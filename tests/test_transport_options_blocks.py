"""The non-default transport options on a decomposed mesh (SURVEY section 8(e) for the rows of tests/test_transport_options.py):
blocks with two halo layers, every block behind its own handle, halo update between steps -- the owned cells must come
out bit-identical to the single-block oracle run (the reference's regression policy across rank counts), and the
conservation sums a host adds over the blocks (ir_set_checks(conservation = 2); the reference's global sum is
mpas_dmpar_sum_real_array, mpas_seaice_advection_incremental_remap.F:8090-8105) must be the single-block sums.

CPU only: the shipped kernels compiled for the host (tests/emu); the same entry points run on a B200 in
tests/test_transport_options.py (one block) and tests/test_ir_blocks.py (blocks, default options)."""
import numpy as np
import pytest

from oracle import ir, upwind
from mpas_seaice_b200 import ir_host, partition, variational_init
from test_oracle_ir import smooth_divergent_velocity, _random_state
from test_ir_parity import _emulation_library, clone
from test_ir_blocks import _blocks, _halo_update
from test_transport_options import _upwind_state, _clone_vars


def _restrict(b, mesh, a):
    return partition.restrict_field(b, a, mesh.nCells, mesh.nVertices)


def _block_upwind_mesh(b, f):
    """interiorEdge and the polygon edge normals of one block, computed on the block like seaice_init_advection_upwind
    does on a rank (mpas_seaice_advection_upwind.F:96-128)."""
    iv = variational_init.interior_vertex(b)
    nve = upwind.normal_vectors(b, f, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    return ir_host.interior_edge(b), nve


@pytest.mark.parametrize("kind,n_parts,table", [("ico4", 3, "physical"), ("hex16", 2, "physical"), ("ico4", 2, "reference")])
def test_upwind_blocks_reproduce_the_single_block_run(kind, n_parts, table):
    """Donor-cell upwind needs the old values of the first halo layer and, for the 'none parent' mask of a child tracer
    (prepare_none_parent_tracer, :566-604), the neighbours of those: two halo layers.  Oracle and emulated kernels on
    every block, three steps, against the single-block oracle."""
    lib = _emulation_library()
    mesh, irf, geom, blocks, birfs = _blocks(kind, n_parts, 2)
    iv = variational_init.interior_vertex(mesh)
    nve = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    interior = ir_host.interior_edge(mesh)
    var = _upwind_state(mesh, np.random.default_rng(23), n_cat=2, table=table)
    u, v = smooth_divergent_velocity(mesh, geom)
    single, gathered_orc, gathered_dev = _clone_vars(var), _clone_vars(var), _clone_vars(var)

    def block_vars():
        return [[upwind.Var(x.name, _restrict(b, mesh, x.array), x.parent, x.volume_like, x.child_minimum) for x in var]
                for b in blocks]

    bvar_orc, bvar_dev = block_vars(), block_vars()
    bmesh = [_block_upwind_mesh(b, f) for b, f in zip(blocks, birfs)]
    buv = [(_restrict(b, mesh, u), _restrict(b, mesh, v)) for b in blocks]
    solvers = []
    try:
        for b, f, (bint, bnve) in zip(blocks, birfs, bmesh):
            g = ir_host.init_geometry(b, f, n_cells_solve=b.nCellsSolve, lib_path=lib)
            s = ir_host.IrTransport(b, f, g, 2, n_cells_solve=b.nCellsSolve, lib_path=lib)
            s.set_upwind_mesh(bint, b.dvEdge, bnve)
            solvers.append(s)
        for _ in range(3):
            upwind.run(mesh, irf["verticesOnEdge"], interior, nve, single, u, v, 3600.0)
            for b, f, (bint, bnve), vo, vd, s, (uu, vv) in zip(blocks, birfs, bmesh, bvar_orc, bvar_dev, solvers, buv):
                upwind.run(b, f["verticesOnEdge"], bint, bnve, vo, uu, vv, 3600.0, n_cells_solve=b.nCellsSolve)
                s.run_upwind(vd, uu, vv, 3600.0)
                nS = b.nCellsSolve
                for x, y in zip(vo, vd):
                    assert np.array_equal(x.array[:nS], y.array[:nS]), x.name
            _halo_update(mesh, blocks, bvar_orc, gathered_orc)
            _halo_update(mesh, blocks, bvar_dev, gathered_dev)
    finally:
        for s in solvers:
            s.destroy()
    nC = mesh.nCells
    assert np.abs(single[0].array[:nC] - var[0].array[:nC]).max() > 1e-6
    for a, go, gd in zip(single, gathered_orc, gathered_dev):
        assert np.array_equal(a.array[:nC], go.array[:nC]), a.name
        assert np.array_equal(a.array[:nC], gd.array[:nC]), a.name


def test_one_halo_layer_is_not_enough_for_a_child_tracer():
    """With one halo layer the 'none parent' mask of the halo cells misses neighbours, and the edge flux of a child
    tracer next to ice-free cells changes: the area (no parent) still agrees, a child does not have to."""
    mesh, irf, geom, blocks, birfs = _blocks("ico4", 3, 1)
    iv = variational_init.interior_vertex(mesh)
    nve = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    var = _upwind_state(mesh, np.random.default_rng(23), n_cat=2, ice_free=0.5)
    u, v = smooth_divergent_velocity(mesh, geom)
    single, gathered = _clone_vars(var), _clone_vars(var)
    upwind.run(mesh, irf["verticesOnEdge"], ir_host.interior_edge(mesh), nve, single, u, v, 3600.0)
    for b, f in zip(blocks, birfs):
        bint, bnve = _block_upwind_mesh(b, f)
        bv = [upwind.Var(x.name, _restrict(b, mesh, x.array), x.parent, x.volume_like, x.child_minimum) for x in var]
        upwind.run(b, f["verticesOnEdge"], bint, bnve, bv, _restrict(b, mesh, u), _restrict(b, mesh, v), 3600.0,
                   n_cells_solve=b.nCellsSolve)
        for t in range(len(bv)):
            partition.scatter_owned(b, bv[t].array, gathered[t].array, "cell")
    nC = mesh.nCells
    assert np.array_equal(single[0].array[:nC], gathered[0].array[:nC])      # the area needs one layer only
    assert np.allclose(single[3].array[:nC], gathered[3].array[:nC], atol=30.0)


@pytest.mark.parametrize("kind,n_parts", [("ico4", 3), ("hex16", 2)])
def test_block_conservation_sums_add_up_to_the_global_sums(kind, n_parts):
    """conservation = 2: every block returns the sums over its owned cells before and after the step; added over the
    blocks they are the single-block sums (1e-13: the order of additions differs), the closed sphere conserves every
    one to the reference's 1e-11 while a single block alone does not, and the transported fields are those of the
    run without checks."""
    lib = _emulation_library()
    mesh, irf, geom, blocks, birfs = _blocks(kind, n_parts, 2)
    tracers = _random_state(mesh, np.random.default_rng(17), n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    single, gathered = clone(tracers), clone(tracers)
    d_ref = ir.run(mesh, irf, geom, single, u, v, 3600.0, conservation_check=2)
    total_init = [np.zeros_like(np.asarray(x, dtype=float)) for x in d_ref["sumInit"]]
    total_final = [np.zeros_like(np.asarray(x, dtype=float)) for x in d_ref["sumFinal"]]
    some_block_is_open = False
    for b, f in zip(blocks, birfs):
        g = ir_host.init_geometry(b, f, n_cells_solve=b.nCellsSolve, lib_path=lib)
        s = ir_host.IrTransport(b, f, g, 2, n_cells_solve=b.nCellsSolve, lib_path=lib)
        try:
            tr = [ir.Tracer(t.name, _restrict(b, mesh, t.array), t.parent, t.volume_like) for t in tracers]
            s.set_tracers(tr)
            s.set_checks(conservation=2, monotonicity=0)
            assert s.run(tr, _restrict(b, mesh, u), _restrict(b, mesh, v), 3600.0) == 0
            for t in range(len(tr)):
                si, sf = s.conservation_sums(t, tr[t].array.shape[2])
                total_init[t] += np.asarray(si).reshape(total_init[t].shape)
                total_final[t] += np.asarray(sf).reshape(total_final[t].shape)
                if np.any(np.abs(np.asarray(sf) - np.asarray(si)) > 1e-11 * np.abs(np.asarray(si))):
                    some_block_is_open = True
                partition.scatter_owned(b, tr[t].array, gathered[t].array, "cell")
        finally:
            s.destroy()
    assert some_block_is_open
    for t in range(len(tracers)):
        ri, rf = np.asarray(d_ref["sumInit"][t], dtype=float), np.asarray(d_ref["sumFinal"][t], dtype=float)
        assert np.all(np.abs(total_init[t] - ri) <= 1e-13 * np.abs(ri).max()), tracers[t].name
        assert np.all(np.abs(total_final[t] - rf) <= 1e-13 * np.abs(rf).max()), tracers[t].name
        if kind == "ico4":       # closed surface: nothing leaves
            assert np.all(np.abs(total_final[t] - total_init[t]) <= 1e-11 * np.abs(total_init[t]).max()), tracers[t].name
        assert np.array_equal(single[t].array[:mesh.nCells], gathered[t].array[:mesh.nCells]), tracers[t].name
