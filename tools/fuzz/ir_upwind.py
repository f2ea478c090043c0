import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
import test_transport_options as T
import test_ir_parity as P
lib = P._emulation_library()
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for kind in ["hex16", "quad16", "ico3", "band48"]:
    mesh, irf, geom, interior, nve = T._upwind_setup(kind)
    nC, nV, nE = mesh.nCells, mesh.nVertices, mesh.nEdges
    for seed in range(lo, hi):
        rng = np.random.default_rng(21000 + seed)
        ncat = int(rng.integers(1, 5))
        table = ["physical", "reference"][seed % 2]
        nCS = int(rng.integers(nC // 3, nC + 1))
        var = T._upwind_state(mesh, rng, n_cat=ncat, table=table, ice_free=rng.uniform(0, 0.8))
        cfl = rng.uniform(0.05, 0.6)
        if seed % 3 == 0:
            speed = cfl * geom["minLengthEdgesOnVertex"][:nV].min() / 3600.0
            u, v = np.zeros(nV + 1), np.zeros(nV + 1)
            u[:nV], v[:nV] = rng.uniform(-speed, speed, nV), rng.uniform(-speed, speed, nV)
        else:
            u, v = T.smooth_divergent_velocity(mesh, geom, cfl=cfl)
        ref, dev = T._clone_vars(var), T._clone_vars(var)
        s = T._solver(kind, lib, ncat, n_cells_solve=nCS)
        try:
            s.set_upwind_mesh(interior, mesh.dvEdge, nve)
            for stepn in range(2):
                d = T.upwind.run(mesh, irf["verticesOnEdge"], interior, nve, ref, u, v, 3600.0, n_cells_solve=nCS, diagnostics=True)
                s.run_upwind(dev, u, v, 3600.0)
                for i, (x, y) in enumerate(zip(ref, dev)):
                    if not np.array_equal(x.array.view(np.int64), y.array.view(np.int64)):
                        bad.append((kind, seed, stepn, x.name)); break
                    flux, vel = s.upwind_fluxes(i)
                    if not np.array_equal(flux[:nE], d["edgeFlux"][i][:nE]):
                        bad.append((kind, seed, stepn, x.name, "flux")); break
                for x, y in zip(ref, dev):
                    y.array[:] = x.array
        except Exception as e:
            bad.append((kind, seed, "EXC", repr(e)[:200]))
        finally:
            s.destroy()
print("upwind seeds", lo, hi, "failures:", bad[:10], len(bad), "%.0fs" % (time.time() - t0))
