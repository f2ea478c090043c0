"""Resident full steps (pre -> subcycles -> post) with a RANDOM ice cover every step: blobs, bands, nothing, everything.
The device keeps u, v, stresses, solveVelocityPrevious between the steps; the oracle chain carries them explicitly."""
import os, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests", os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from mpas_seaice_b200 import host, variational_init, synthetic
import evp_emu
host._lib = host.load_library(evp_emu.library())
import common, oracle
import test_gpu_prepost as P

def cover(rng, mesh):
    nC = mesh.nCells
    mode = rng.integers(0, 5)
    u = rng.uniform(size=nC)
    if mode == 0:
        on = u < rng.uniform(0.05, 0.95)
    elif mode == 1:
        x = mesh.latCell[:nC] if mesh.on_a_sphere else mesh.xCell[:nC] / mesh.xCell[:nC].max()
        t = rng.uniform(x.min(), x.max())
        on = x > t if rng.uniform() < 0.5 else x < t
    elif mode == 2:
        lo = rng.integers(0, nC); hi = min(nC, lo + rng.integers(1, nC))
        on = np.zeros(nC, bool); on[lo:hi] = True
    elif mode == 3:
        on = np.ones(nC, bool)
    else:
        on = np.zeros(nC, bool)
        on[rng.integers(0, nC, size=rng.integers(1, 6))] = True
    area = np.where(on, rng.uniform(0.2, 1.0, nC), 0.0)
    vol = np.where(on, area * rng.uniform(0.2, 3.0, nC), 0.0)
    return area, vol

bad = []
t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(7000 + seed)
    kind = ["hex20", "ico3", "quad40", "ico4"][seed % 4]
    mesh, var = common.mesh_case(kind)
    base = P._state(mesh, "square" if not mesh.on_a_sphere else "B")
    nC, nV = mesh.nCells, mesh.nVertices
    interior = variational_init.interior_vertex(mesh)
    cr = str(rng.choice(["evp", "evp_revised"]))
    _, opts = synthetic.pre_subcycle(mesh, base, 3600.0, constitutive_relation_type=cr)
    opts = dict(opts, ocean_stress_type=str(rng.choice(["quadratic", "linear"])))
    solver = P._solver(mesh, var, opts)
    prev = None
    try:
        for it in range(5):
            state = dict(base)
            area, vol = cover(rng, mesh)
            for k, a in (("iceAreaCell", area), ("iceVolumeCell", vol), ("snowVolumeCell", 0.1 * vol)):
                z = np.zeros(nC + 1); z[:nC] = a; state[k] = z
            n_sub = int(rng.integers(1, 8))
            ref_step = oracle.pre_subcycle(mesh, state, 3600.0, prev=prev)
            solver.pre_subcycle(P._cells(mesh, state), cold_start=(it == 0))
            got_pre = solver.fetch_pre()
            vm = ref_step["solveVelocity"][:nV] == 1
            try:
                assert np.array_equal(got_pre["solveStress"][:nC], ref_step["solveStress"][:nC]), "solveStress"
                assert np.array_equal(got_pre["solveVelocity"][:nV], ref_step["solveVelocity"][:nV]), "solveVelocity"
                assert np.array_equal(got_pre["solveVelocityPrevious"][:nV], ref_step["solveVelocityPrevious"][:nV]), "prev"
                oracle.subcycle_velocity_solver(mesh, var, ref_step, opts, n_sub)
                solver.run_subcycles(n_sub)
                ref = P._post_reference(mesh, ref_step, opts, interior)
                got = solver.post_subcycle(names=host.POST_FIELDS_VARIATIONAL)
                inner = solver.fetch(names=("stress11", "stress22", "stress12", "uVelocity", "vVelocity"))
                cmc = ref_step["solveStress"][:nC] == 1
                for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV"):
                    assert np.array_equal(got[k][:nC], ref[k][:nC]), k
                for k in ("uVelocity", "vVelocity"):
                    assert np.array_equal(got[k][:nV], ref[k][:nV]), k + " (all vertices)"
                for k in ("oceanStressU", "oceanStressV", "oceanStressCoeff"):
                    assert np.array_equal(got[k][:nV][vm], ref[k][:nV][vm]), k
                valid = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
                for k in ("stress11", "stress22", "stress12"):
                    assert np.array_equal(inner[k][:nC][valid], ref_step[k][:nC][valid]), k + " (all cells)"
            except AssertionError as e:
                bad.append((seed, it, str(e)[:100])); break
            prev = dict(uVelocity=ref_step["uVelocity"], vVelocity=ref_step["vVelocity"], stress11=ref_step["stress11"],
                        stress22=ref_step["stress22"], stress12=ref_step["stress12"],
                        solveVelocityPrevious=ref_step["solveVelocityPrevious"])
    except Exception as e:
        bad.append((seed, "EXC", repr(e)[:300]))
    finally:
        solver.destroy()
print("seeds", lo, hi, "persistent=%s" % os.environ.get("EVP_B200_PERSISTENT", "default"), "failures:", bad, "%.0fs" % (time.time() - t0))
