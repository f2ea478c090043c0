"""More seeds of the IR fuzz on the emulated kernels, with category / layer counts varied and checks on."""
import os, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
import test_ir_parity as T
from test_ir_parity import *   # case, _random_state, clone, smooth_divergent_velocity, ir, ir_host
lib = T._emulation_library()
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
aborts = 0; many = 0
for kind in ["hex16", "quad16", "ico3", "band48"]:
    mesh, irf, geom = case(kind)
    nC, nV = mesh.nCells, mesh.nVertices
    for (nk, ni, ns) in [(1, 1, 0), (3, 3, 2), (2, 1, 1)]:
        solver = ir_host.IrTransport(mesh, irf, geom, nk, lib_path=lib)
        try:
            for seed in range(lo, hi):
                rng = np.random.default_rng(50000 + seed * 7 + nk)
                tracers = T._random_state(mesh, rng, n_cat=nk, n_ice=ni, n_snow=ns, ice_free=rng.uniform(0, 0.8))
                cfl = rng.uniform(0.05, 0.7)
                if seed % 2 == 0:
                    speed = cfl * geom["minLengthEdgesOnVertex"][:nV].min() / 3600.0
                    u, v = np.zeros(nV + 1), np.zeros(nV + 1)
                    u[:nV], v[:nV] = rng.uniform(-speed, speed, nV), rng.uniform(-speed, speed, nV)
                else:
                    uu, vv = smooth_divergent_velocity(mesh, geom, cfl=cfl)
                    ang = rng.uniform(0, 2 * np.pi)
                    u, v = uu * np.cos(ang) - vv * np.sin(ang), uu * np.sin(ang) + vv * np.cos(ang)
                ref, dev = clone(tracers), clone(tracers)
                d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, check=False, diagnostics=True)
                solver.set_tracers(dev)
                rc = solver.run(dev, u, v, 3600.0, check=False)
                d_dev = solver.diagnostics()
                try:
                    for key in d_dev:
                        assert np.array_equal(d_ref[key], d_dev[key]), (seed, key)
                    assert (d_ref["error"] != 0) == (rc != 0), (seed, d_ref["error"], rc)
                    if d_ref["error"] == 0:
                        for a, b in zip(ref, dev):
                            assert np.array_equal(a.array[:nC], b.array[:nC]), (seed, a.name)
                    else:
                        aborts += 1
                    many += int(np.count_nonzero(d_ref["triangleArea"], axis=1).max() > 4)
                except AssertionError as e:
                    bad.append((kind, nk, ni, ns, str(e)[:120]))
        finally:
            solver.destroy()
print("seeds", lo, hi, "failures:", bad, "aborting steps:", aborts, "steps with > 4 triangles:", many, "%.0fs" % (time.time() - t0))
