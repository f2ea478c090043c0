import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from oracle import ir
from mpas_seaice_b200 import partition, ir_host
from test_oracle_ir import case, smooth_divergent_velocity, _random_state
import test_ir_parity as T
from test_ir_parity import clone
import test_ir_blocks as B
lib = T._emulation_library()
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(41000 + seed)
    kind = ["hex16", "ico3", "quad16"][seed % 3]
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    P = int(rng.integers(2, 5))
    mode = seed % 3
    if mode == 0:
        part = partition.partition_cells(mesh, P).copy()
        flip = rng.uniform(size=nC) < 0.08
        part[flip] = rng.integers(0, P, int(flip.sum()))
    elif mode == 1:
        part = np.searchsorted(np.sort(rng.integers(1, nC - 1, P - 1)), np.arange(nC), side="right")
    else:
        part = rng.integers(0, P, nC)
    part = np.asarray(part, dtype=np.int64)
    solvers = []
    try:
        nh = 2 if int(mesh.vertexDegree) == 3 else 3
        blocks = [partition.build_block(mesh, part, r, nh) for r in range(P)]
        birfs = [partition.restrict_ir(b, mesh, irf) for b in blocks]
        ncat = int(rng.integers(1, 3))
        tracers = _random_state(mesh, rng, n_cat=ncat, n_ice=int(rng.integers(1, 3)), n_snow=int(rng.integers(0, 2)))
        u, v = smooth_divergent_velocity(mesh, geom, cfl=rng.uniform(0.1, 0.5))
        single, gathered = clone(tracers), clone(tracers)
        btr, buv, live = [], [], []
        for b, f in zip(blocks, birfs):
            ok = b.nCellsSolve > 0
            live.append(ok)
            tr = [ir.Tracer(t.name, partition.restrict_field(b, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like) for t in tracers]
            btr.append(tr)
            buv.append((partition.restrict_field(b, u, mesh.nCells, mesh.nVertices), partition.restrict_field(b, v, mesh.nCells, mesh.nVertices)))
            if ok:
                g = ir_host.init_geometry(b, f, n_cells_solve=b.nCellsSolve, lib_path=lib)
                s = ir_host.IrTransport(b, f, g, ncat, n_cells_solve=b.nCellsSolve, lib_path=lib)
                s.set_tracers(tr)
                solvers.append(s)
            else:
                solvers.append(None)
        for _ in range(2):
            ir.run(mesh, irf, geom, single, u, v, 3600.0)
            for b, s, tr, (uu, vv), ok in zip(blocks, solvers, btr, buv, live):
                if ok:
                    s.run(tr, uu, vv, 3600.0)
            B._halo_update(mesh, blocks, btr, gathered)
        for a, g in zip(single, gathered):
            assert np.array_equal(a.array[:nC], g.array[:nC]), a.name
    except AssertionError as e:
        bad.append((seed, kind, P, mode, "ASSERT " + str(e)[:80]))
    except Exception as e:
        import traceback
        bad.append((seed, kind, P, mode, traceback.format_exc()[-300:]))
    finally:
        for s in solvers:
            if s is not None:
                s.destroy()
print("ir device-block seeds", lo, hi, "failures:", len(bad), bad[:4], "%.0fs" % (time.time() - t0))
