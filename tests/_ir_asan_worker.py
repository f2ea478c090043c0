"""Drives the AddressSanitizer build of the emulated IR kernels (argv[1]) through geometry init, the full tracer
hierarchy on hexes / quads / the sphere, and decomposed blocks with halo cells.  Launched by
tests/test_ir_parity.py::test_emulated_kernels_are_clean_under_address_sanitizer with libasan preloaded: the
emulated device memory is plain calloc memory of exactly the requested size, so an index past the end of a device
array is reported and aborts this process."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import ir
from mpas_seaice_b200 import ir_host, partition
from test_oracle_ir import case, smooth_divergent_velocity, _random_state, uniform_velocity
lib = sys.argv[1]
for kind in ("hex16", "quad16", "ico3", "band48"):
    mesh, irf, _ = case(kind)
    geom = ir_host.init_geometry(mesh, irf, lib_path=lib)
    rng = np.random.default_rng(21)
    tr = _random_state(mesh, rng)
    u, v = smooth_divergent_velocity(mesh, geom)
    s = ir_host.IrTransport(mesh, irf, geom, 3, lib_path=lib)
    s.set_tracers(tr)
    for _ in range(2): s.run(tr, u, v, 3600.0)
    d = s.diagnostics()
    s.destroy()
    print(kind, "ok")
# blocks with halo
mesh, irf, _ = case("ico4")
part = partition.partition_cells(mesh, 3)
for r in range(3):
    b = partition.build_block(mesh, part, r, 2)
    f = partition.restrict_ir(b, mesh, irf)
    g = ir_host.init_geometry(b, f, n_cells_solve=b.nCellsSolve, lib_path=lib)
    rng = np.random.default_rng(17)
    tr = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    btr = [ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like) for t in tr]
    u, v = smooth_divergent_velocity(mesh, case("ico4")[2])
    s = ir_host.IrTransport(b, f, g, 2, n_cells_solve=b.nCellsSolve, lib_path=lib)
    s.set_tracers(btr)
    s.run(btr, partition.restrict_field(b, u, mesh.nCells, mesh.nVertices), partition.restrict_field(b, v, mesh.nCells, mesh.nVertices), 3600.0)
    s.destroy()
    print("block", r, "ok")
# the non-default options: checks, normal vectors, upwind (tests/test_transport_options.py)
from oracle import upwind
from mpas_seaice_b200 import variational_init
from test_transport_options import _upwind_state
for kind in ("hex16", "band48"):
    mesh, irf, geom = case(kind)
    nCS = (2 * mesh.nCells) // 3
    tr = _random_state(mesh, np.random.default_rng(21))
    u, v = smooth_divergent_velocity(mesh, geom)
    s = ir_host.IrTransport(mesh, irf, geom, 3, n_cells_solve=nCS, lib_path=lib)
    s.set_tracers(tr)
    s.set_checks(2, 1)
    s.run(tr, u, v, 3600.0, check=False)
    iv = variational_init.interior_vertex(mesh)
    for rm in (True, False):
        nv = ir_host.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=rm, lib_path=lib)
    var = _upwind_state(mesh, np.random.default_rng(3))
    s.set_upwind_mesh(ir_host.interior_edge(mesh), mesh.dvEdge, nv["normalVectorPolygon"])
    for _ in range(2): s.run_upwind(var, u, v, 3600.0)
    s.upwind_fluxes(3)
    s.destroy()
    print(kind, "options ok")
