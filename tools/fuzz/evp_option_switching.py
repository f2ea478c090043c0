"""Options switched on ONE handle between steps (evp_set_options): constitutive relation, drag, ocean stress on / off,
averaged strains, the weak schemes (with the weak mesh set once), special boundaries -- each step against the oracle."""
import os, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests", os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from mpas_seaice_b200 import host, variational_init, weakmesh
import evp_emu
host._lib = host.load_library(evp_emu.library())
import common
from test_gpu_parity import _compare
import test_gpu_weak as W

bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(11000 + seed)
    kind = ["hex20", "quad40", "ico3", "ico4"][seed % 4]
    mesh, var = common.mesh_case(kind)
    weak = weakmesh.weak_fields(mesh)
    state = "auto" if not mesh.on_a_sphere else str(rng.choice(["A", "B"]))
    base, opts0 = common.step_case(mesh, state_kind=state)
    nC, nV = mesh.nCells, mesh.nVertices
    solver = host.EvpSolver(mesh, var, opts0)
    try:
        solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.set_weak_mesh(mesh, weak)
        for it in range(6):
            cr = str(rng.choice(["evp", "evp_revised", "linear", "none"]))
            _, o = common.step_case(mesh, state_kind=state, constitutive_relation_type=cr)
            scheme = [("variational", "variational"), ("weak", "weak"), ("weak", "variational")][int(rng.integers(0, 3))]
            opts = dict(o, ocean_stress_type=str(rng.choice(["quadratic", "linear"])), use_ocean_stress=bool(rng.uniform() < 0.8),
                        average_variational_strain=bool(scheme[0] == "variational" and rng.uniform() < 0.3),
                        strain_scheme=scheme[0], stress_divergence_scheme=scheme[1])
            step = common.clone_step(base)
            step["solveStress"][:nC][rng.uniform(size=nC) < rng.uniform(0, 0.5)] = 0
            step["solveVelocity"][:nV][rng.uniform(size=nV) < rng.uniform(0, 0.5)] = 0
            on_v = step["solveVelocity"] == 1
            step["uVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            step["vVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
            step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
            on_c = (step["solveStress"] == 1)
            for k in ("stress11", "stress22", "stress12"):
                step[k] = np.where(on_c[:, None], rng.uniform(-500.0, 500.0, step[k].shape), 0.0)
            if scheme[0] == "weak":
                for k in ("stress11Weak", "stress22Weak", "stress12Weak"):
                    step[k] = np.where(on_c, rng.uniform(-500.0, 500.0, nC + 1), 0.0)
            n_sub = int(rng.integers(1, 6))
            ref = common.run_oracle(mesh, dict(var, weak=weak), step, opts, n_sub)
            solver.set_options(opts)
            solver.update_step(step)
            if scheme[0] == "weak":
                solver.update_weak_state({k: step[k] for k in ("stress11Weak", "stress22Weak", "stress12Weak")})
            solver.run_subcycles(n_sub)
            out = solver.fetch()
            try:
                if scheme == ("weak", "weak"):
                    wk = solver.fetch_weak()
                    _, vm = common.masks_for(mesh, step)
                    for k in ("uVelocity", "vVelocity"):
                        assert np.array_equal(out[k][vm], ref[k][vm]), k
                    for k in ("stress11Weak", "stress22Weak", "stress12Weak"):
                        assert np.array_equal(wk[k][:nC], ref[k][:nC]), k
                else:
                    _compare(mesh, step, ref, out)
            except AssertionError as e:
                bad.append((seed, it, cr, scheme, str(e)[:100])); break
    except Exception as e:
        import traceback
        bad.append((seed, "EXC", traceback.format_exc()[-400:]))
    finally:
        solver.destroy()
print("seeds", lo, hi, "failures:", bad, "%.0fs" % (time.time() - t0))
