"""pre-subcycle switches, ice shelves and several categories at random, whole resident steps, emulated device vs oracle."""
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
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(61000 + seed)
    kind = ["hex20", "ico3", "quad40", "ico4"][seed % 4]
    mesh, var = common.mesh_case(kind)
    base = P._state(mesh, "B" if mesh.on_a_sphere else "square")
    nC, nV = mesh.nCells, mesh.nVertices
    interior = variational_init.interior_vertex(mesh)
    _, opts = synthetic.pre_subcycle(mesh, base, 3600.0)
    sw = dict(use_air_stress=bool(rng.uniform() < 0.7), use_surface_tilt=bool(rng.uniform() < 0.7),
              geostrophic_surface_tilt=bool(rng.uniform() < 0.6))
    land = None
    if rng.uniform() < 0.5:
        land = np.zeros(nC + 1, np.int32)
        land[:nC] = rng.uniform(size=nC) < rng.uniform(0.05, 0.4)
    land_vertex = variational_init.land_ice_mask_vertex(mesh, land) if land is not None else None
    ncat = int(rng.integers(1, 4))
    solver = host.EvpSolver(mesh, var, opts)
    solver.set_mesh_ext(mesh, interior, **({"land_ice_mask_vertex": land_vertex} if land is not None else {}))
    M = mesh.maxEdges
    prev = dict(uVelocity=np.zeros(nV + 1), vVelocity=np.zeros(nV + 1), solveVelocityPrevious=np.zeros(nV + 1, dtype=np.int32),
                stress11=np.zeros((nC + 1, M)), stress22=np.zeros((nC + 1, M)), stress12=np.zeros((nC + 1, M)))
    try:
        for it in range(3):
            w = rng.uniform(0.1, 1.0, ncat); w /= w.sum()
            on = rng.uniform(size=nC) < rng.uniform(0.2, 1.0)
            area = np.where(on, rng.uniform(0.2, 1.0, nC), 0.0)
            a = np.zeros((nC + 1, ncat)); vi = np.zeros((nC + 1, ncat)); vs = np.zeros((nC + 1, ncat))
            for k in range(ncat):
                a[:nC, k] = area * w[k]; vi[:nC, k] = area * w[k] * rng.uniform(0.3, 3.0, nC); vs[:nC, k] = 0.1 * vi[:nC, k]
            A, VI, VS, mass = oracle.aggregate_mass_and_area(a, vi, vs)
            state = dict(base, iceAreaCell=A, iceVolumeCell=VI, snowVolumeCell=VS)
            forcing = {}
            if not sw["geostrophic_surface_tilt"]:
                forcing = dict(seaSurfaceTiltU=1e-6 * rng.uniform(-1, 1, nC + 1), seaSurfaceTiltV=1e-6 * rng.uniform(-1, 1, nC + 1))
                state.update(forcing)
            kw = dict(sw)
            if land is not None:
                kw.update(land_ice_mask=land, land_ice_mask_vertex=land_vertex)
            ref_step = oracle.pre_subcycle(mesh, state, 3600.0, prev=prev, **kw)
            solver.aggregate(a.copy(), vi.copy(), vs.copy(), hibler_strength=False)
            agg = solver.fetch_aggregate(ice_pressure=False)
            try:
                for k, want in (("iceAreaCell", A), ("iceVolumeCell", VI), ("snowVolumeCell", VS), ("totalMassCell", mass)):
                    assert np.array_equal(agg[k][:nC], want[:nC]), "agg " + k
                cells = dict({k: np.ascontiguousarray(state[k], dtype=np.float64) for k in ("uOceanVelocity", "vOceanVelocity", "uAirVelocity", "vAirVelocity", "airDensity")},
                             iceAreaCellInitial=agg["iceAreaCell"], iceAreaCell=agg["iceAreaCell"], totalMassCell=agg["totalMassCell"],
                             icePressure=oracle.hibler_strength_unmasked(state, nC), **forcing)
                if land is not None:
                    cells["landIceMask"] = land
                n_sub = int(rng.integers(1, 6))
                solver.pre_subcycle(cells, cold_start=(host.START_FIRST_STEP if it == 0 else host.START_RESIDENT), **sw)
                got_pre = solver.fetch_pre()
                for k, n in (("solveStress", nC), ("solveVelocity", nV), ("solveVelocityPrevious", nV)):
                    assert np.array_equal(got_pre[k][:n], ref_step[k][:n]), "pre " + k
                vm = ref_step["solveVelocity"][:nV] == 1
                for k in ("airStressVertexU", "surfaceTiltForceU", "surfaceTiltForceV", "oceanStressU", "totalMassVertexfVertex"):
                    assert np.array_equal(got_pre[k][:nV][vm], ref_step[k][:nV][vm]), "pre " + k
                oracle.subcycle_velocity_solver(mesh, var, ref_step, opts, n_sub)
                solver.run_subcycles(n_sub)
                ref = P._post_reference(mesh, ref_step, opts, interior)
                got = solver.post_subcycle(names=host.POST_FIELDS_VARIATIONAL)
                for k in ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV"):
                    assert np.array_equal(got[k][:nC], ref[k][:nC]), k
                for k in ("uVelocity", "vVelocity"):
                    assert np.array_equal(got[k][:nV], ref[k][:nV]), k
            except AssertionError as e:
                bad.append((seed, it, sw, land is not None, ncat, str(e)[:80])); break
            prev = {k: ref_step[k] for k in ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "solveVelocityPrevious")}
    except Exception as e:
        import traceback
        bad.append((seed, "EXC", traceback.format_exc()[-400:]))
    finally:
        solver.destroy()
print("seeds", lo, hi, "failures:", len(bad), bad[:4], "%.0fs" % (time.time() - t0))
