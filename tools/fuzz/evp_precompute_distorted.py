"""Wachspress / PWL precompute on DISTORTED cells (vertices jittered), every quadrature rule: emulated device vs oracle."""
import os, sys, os, time, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests", os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from mpas_seaice_b200 import host, meshgen
import evp_emu
host._lib = host.load_library(evp_emu.library())
import common, oracle
RULES = [("dunavant", o) for o in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)] + [("fekete", o) for o in (1, 2, 3, 4, 5, 6, 8, 9)] + [("trapezoidal", o) for o in (1, 2, 3, 5)]
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(13000 + seed)
    which = seed % 3
    if which == 0:
        mesh = meshgen.planar_hex(8, 9, 16000.0)
    elif which == 1:
        mesh = meshgen.planar_quad(7, 7, 16000.0)
    else:
        mesh = meshgen.icosphere(2)
    mesh = meshgen.Mesh(mesh)
    nV = mesh.nVertices
    amp = rng.uniform(0.0, 0.22)
    if mesh.on_a_sphere:
        dc = float(mesh.dcEdge[:-1].mean())
        p = np.stack([mesh.xVertex[:nV], mesh.yVertex[:nV], mesh.zVertex[:nV]], 1) + amp * dc * rng.uniform(-1, 1, (nV, 3))
        p *= (mesh.sphere_radius / np.linalg.norm(p, axis=1))[:, None]
        mesh.xVertex = mesh.xVertex.copy(); mesh.yVertex = mesh.yVertex.copy(); mesh.zVertex = mesh.zVertex.copy()
        mesh.xVertex[:nV], mesh.yVertex[:nV], mesh.zVertex[:nV] = p[:, 0], p[:, 1], p[:, 2]
    else:
        mesh.xVertex = mesh.xVertex.copy(); mesh.yVertex = mesh.yVertex.copy()
        mesh.xVertex[:nV] += amp * 16000.0 * rng.uniform(-1, 1, nV)
        mesh.yVertex[:nV] += amp * 16000.0 * rng.uniform(-1, 1, nV)
    basis = "pwl" if rng.uniform() < 0.2 else "wachspress"
    itype, order = RULES[int(rng.integers(0, len(RULES)))]
    try:
        kw = dict(basis=basis) if basis == "pwl" else dict(integration_type=itype, integration_order=order)
        var = oracle.init_variational(mesh, **kw)
        step, opts = common.step_case(mesh)
        solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]), integration=(itype, order), basis=basis)
        try:
            got = solver.fetch_basis()
        finally:
            solver.destroy()
        for k, a in got.items():
            if not np.array_equal(a[:mesh.nCells], var[k][:mesh.nCells]):
                d = np.abs(a[:mesh.nCells] - var[k][:mesh.nCells]).max()
                bad.append((seed, basis, itype, order, k, float(d))); break
    except Exception as e:
        bad.append((seed, basis, itype, order, "EXC", repr(e)[:200]))
print("seeds", lo, hi, "failures:", bad, "%.0fs" % (time.time() - t0))
print("sanity: last case", basis, itype, order, "amp", amp, "max|GU|", float(np.abs(got["basisGradientU"]).max()), "nonzero SU", int(np.count_nonzero(got["basisIntegralsU"])))
