"""Sequences of steps on ONE handle: masks, forcing, state and subcycle counts change from step to step (the device keeps its
contrib rows, tile lists and graph between them); every step compared with the oracle bit for bit."""
import os, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests", os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from mpas_seaice_b200 import host, variational_init
import evp_emu
host._lib = host.load_library(evp_emu.library())
import common
from test_gpu_parity import _compare

def perturb(rng, mesh, base, frac_hi):
    step = common.clone_step(base)
    nC, nV = mesh.nCells, mesh.nVertices
    mode = rng.integers(0, 4)
    if mode == 0:      # random holes
        step["solveStress"][:nC][rng.uniform(size=nC) < rng.uniform(0, frac_hi)] = 0
        step["solveVelocity"][:nV][rng.uniform(size=nV) < rng.uniform(0, frac_hi)] = 0
    elif mode == 1:    # a contiguous band of cells only (whole tiles without work)
        lo = rng.integers(0, nC); hi = min(nC, lo + rng.integers(1, nC))
        keep = np.zeros(nC, bool); keep[lo:hi] = True
        step["solveStress"][:nC][~keep] = 0
        lo = rng.integers(0, nV); hi = min(nV, lo + rng.integers(1, nV))
        keepv = np.zeros(nV, bool); keepv[lo:hi] = True
        step["solveVelocity"][:nV][~keepv] = 0
    elif mode == 2:    # nothing at all / everything
        if rng.uniform() < 0.3:
            step["solveStress"][:] = 0
        if rng.uniform() < 0.3:
            step["solveVelocity"][:] = 0
    on_v = step["solveVelocity"] == 1
    step["uVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
    step["vVelocity"] = np.where(on_v, rng.uniform(-0.2, 0.2, nV + 1), 0.0)
    if "uVelocityInitial" in step:
        step["uVelocityInitial"], step["vVelocityInitial"] = step["uVelocity"].copy(), step["vVelocity"].copy()
    on_c = (step["solveStress"] == 1)[:, None]
    for k in ("stress11", "stress22", "stress12"):
        step[k] = np.where(on_c, rng.uniform(-500.0, 500.0, step[k].shape), 0.0)
    return step

bad = []
t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(9000 + seed)
    kind = ["hex20", "quad40", "ico3", "ico4"][seed % 4]
    mesh, var = common.mesh_case(kind)
    cr = str(rng.choice(["evp", "evp_revised"]))
    state = "auto" if kind.startswith(("hex", "quad")) else str(rng.choice(["A", "B"]))
    base, opts = common.step_case(mesh, state_kind=state, constitutive_relation_type=cr)
    opts = dict(opts, ocean_stress_type=str(rng.choice(["quadratic", "linear"])), average_variational_strain=bool(rng.uniform() < 0.2))
    solver = host.EvpSolver(mesh, var, opts)
    try:
        if opts.get("average_variational_strain"):
            solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        for n in range(5):
            step = perturb(rng, mesh, base, 0.9)
            n_sub = int(rng.integers(1, 7))
            ref = common.run_oracle(mesh, var, step, opts, n_sub)
            solver.update_step(step)
            if rng.uniform() < 0.5:           # split the run: the second call continues from the resident state
                k = int(rng.integers(0, n_sub + 1))
                if k: solver.run_subcycles(k)
                if n_sub - k: solver.run_subcycles(n_sub - k)
            else:
                solver.run_subcycles(n_sub)
            out = solver.fetch()
            try:
                _compare(mesh, step, ref, out)
            except AssertionError as e:
                bad.append((seed, n, str(e)[:120])); break
    except Exception as e:
        bad.append((seed, "EXC", repr(e)[:200]))
    finally:
        solver.destroy()
print("seeds", lo, hi, "persistent=%s" % os.environ.get("EVP_B200_PERSISTENT", "default"), "failures:", bad, "%.0fs" % (time.time() - t0))
