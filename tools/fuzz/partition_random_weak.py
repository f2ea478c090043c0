import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
import common, oracle
from mpas_seaice_b200 import partition, weakmesh
bad = []; t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(51000 + seed)
    kind = ["hex20", "ico3", "quad40"][seed % 3]
    mesh, var = common.mesh_case(kind)
    gweak = weakmesh.weak_fields(mesh)
    nC, nV = mesh.nCells, mesh.nVertices
    P = int(rng.integers(2, 6))
    mode = seed % 3
    if mode == 0:
        part = rng.integers(0, P, nC)
    elif mode == 1:
        part = partition.partition_cells(mesh, P, "rcb").copy()
        flip = rng.uniform(size=nC) < 0.1
        part[flip] = rng.integers(0, P, int(flip.sum()))
    else:
        part = np.searchsorted(np.sort(rng.integers(1, nC - 1, P - 1)), np.arange(nC), side="right")
    part = np.asarray(part, dtype=np.int64)
    schemes = [("weak", "weak"), ("weak", "variational")][seed % 2]
    step, opts = common.step_case(mesh)
    opts = dict(opts, strain_scheme=schemes[0], stress_divergence_scheme=schemes[1])
    step["solveStress"][:nC][rng.uniform(size=nC) < 0.2] = 0
    step["solveVelocity"][:nV][rng.uniform(size=nV) < 0.2] = 0
    nsub = int(rng.integers(2, 6))
    try:
        ref = common.run_oracle(mesh, dict(var, weak=gweak), step, opts, nsub)
        NH = (int(sys.argv[3]) if len(sys.argv) > 3 else 2) if schemes[1] == "variational" else None
        blocks = [partition.build_block(mesh, part, r, (NH + (1 if int(mesh.vertexDegree) == 4 else 0)) if NH else None) for r in range(P)]
        requests = {r: partition.halo_requests(b) for r, b in enumerate(blocks)}
        lists = [partition.exchange_lists(b, requests) for b in blocks]
        bsteps = [partition.restrict_step(b, step, nC, nV) for b in blocks]
        bvars = []
        for b in blocks:
            if b.nCells == 0:
                bvars.append(None); continue
            v = oracle.init_variational(b)
            v["weak"] = partition.restrict_weak(b, mesh, gweak)
            bvars.append(v)
        for _ in range(nsub):
            for b, v, s in zip(blocks, bvars, bsteps):
                if v is not None:
                    oracle.subcycle_velocity_solver(b, v, s, dict(opts, nVerticesSolve=int(b.nVerticesSolve)), 1)
            common.exchange_halos(bsteps, lists)
        for b, s, v in zip(blocks, bsteps, bvars):
            if v is None: continue
            nVs, nCs = int(b.nVerticesSolve), int(b.nCellsSolve)
            gv = b.indexToVertexID[:nVs].astype(np.int64) - 1
            gc = b.indexToCellID[:nCs].astype(np.int64) - 1
            vm = step["solveVelocity"][gv] == 1
            for k in ("uVelocity", "vVelocity"):
                assert np.array_equal(s[k][:nVs][vm], ref[k][gv][vm]), k
            cmk = step["solveStress"][gc] == 1
            if schemes == ("weak", "weak"):
                for k in ("stress11Weak", "stress22Weak", "stress12Weak"):
                    assert np.array_equal(s[k][:nCs][cmk], ref[k][gc][cmk]), k
            else:
                for k in ("stress11", "stress12"):
                    assert np.array_equal(s[k][:nCs][cmk], ref[k][gc][cmk]), k
    except AssertionError as e:
        bad.append((seed, kind, P, mode, schemes, "ASSERT " + str(e)[:80]))
    except Exception as e:
        import traceback
        bad.append((seed, kind, P, mode, schemes, traceback.format_exc()[-300:]))
print("weak partition seeds", lo, hi, "failures:", len(bad), bad[:5], "%.0fs" % (time.time() - t0))
