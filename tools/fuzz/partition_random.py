"""Random cell-to-rank assignments (scattered cells, tiny parts, an empty part): blocks + halo lists + decomposed run must
reproduce the single-block oracle bit for bit."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
import common, oracle
from mpas_seaice_b200 import partition
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(31000 + seed)
    kind = ["hex20", "ico3", "quad40"][seed % 3]
    mesh, var = common.mesh_case(kind)
    nC, nV = mesh.nCells, mesh.nVertices
    P = int(rng.integers(2, 7))
    mode = seed % 4
    if mode == 0:        # every cell at random
        part = rng.integers(0, P, nC)
    elif mode == 1:      # a proper partition with random cells reassigned
        part = partition.partition_cells(mesh, P, "rcb" ).copy()
        flip = rng.uniform(size=nC) < 0.1
        part[flip] = rng.integers(0, P, int(flip.sum()))
    elif mode == 2:      # one tiny part
        part = partition.partition_cells(mesh, P - 1, "block").copy() if P > 2 else np.zeros(nC, int)
        part[rng.integers(0, nC, 2)] = P - 1
    else:                # contiguous index runs of random lengths
        cuts = np.sort(rng.integers(0, nC, P - 1))
        part = np.searchsorted(cuts, np.arange(nC), side="right")
    part = np.asarray(part, dtype=np.int64)
    step, opts = common.step_case(mesh)
    step["solveStress"][:nC][rng.uniform(size=nC) < 0.2] = 0
    step["solveVelocity"][:nV][rng.uniform(size=nV) < 0.2] = 0
    n_sub = int(rng.integers(2, 6))
    try:
        ref = common.run_oracle(mesh, var, step, opts, n_sub)
        blocks = [partition.build_block(mesh, part, r, None) for r in range(P)]
        requests = {r: partition.halo_requests(b) for r, b in enumerate(blocks)}
        lists = [partition.exchange_lists(b, requests) for b in blocks]
        bvars = [oracle.init_variational(b) if b.nCells > 0 else None for b in blocks]
        bsteps = [partition.restrict_step(b, step, nC, nV) for b in blocks]
        bopts = [dict(opts, nVerticesSolve=int(b.nVerticesSolve)) for b in blocks]
        for _ in range(n_sub):
            for b, v, s, o in zip(blocks, bvars, bsteps, bopts):
                if b.nCells > 0:
                    oracle.subcycle_velocity_solver(b, v, s, o, 1)
            common.exchange_halos(bsteps, lists)
        out = {k: np.zeros_like(step[k]) for k in common.COMPARE_CELL + common.COMPARE_VERTEX}
        for b, s in zip(blocks, bsteps):
            for k in common.COMPARE_CELL:
                partition.scatter_owned(b, s[k], out[k], "cell")
            for k in common.COMPARE_VERTEX:
                partition.scatter_owned(b, s[k], out[k], "vertex")
        cm, vm = common.masks_for(mesh, step)
        for k in common.COMPARE_CELL:
            assert np.array_equal(out[k][cm], ref[k][cm]), k
        for k in common.COMPARE_VERTEX:
            assert np.array_equal(out[k][vm], ref[k][vm]), k
    except AssertionError as e:
        bad.append((seed, kind, P, mode, "ASSERT " + str(e)[:80]))
    except Exception as e:
        import traceback
        bad.append((seed, kind, P, mode, traceback.format_exc()[-300:]))
print("partition seeds", lo, hi, "failures:", len(bad), bad[:6], "%.0fs" % (time.time() - t0))
