"""Shared case builders for the tests: mesh + oracle precompute + synthetic per-step fields."""
from __future__ import annotations

import copy
import functools

import numpy as np

import oracle
from mpas_seaice_b200 import meshgen, synthetic

COMPARE_CELL = ("stress11", "stress22", "stress12", "strain11", "strain22", "strain12", "replacementPressure")
COMPARE_VERTEX = ("uVelocity", "vVelocity", "stressDivergenceU", "stressDivergenceV", "oceanStressCoeff")
POST_CELL = ("divergence", "shear", "ridgeConvergence", "ridgeShear", "oceanStressCellU", "oceanStressCellV")
POST_VERTEX = ("oceanStressU", "oceanStressV")


@functools.lru_cache(maxsize=None)
def mesh_case(kind: str, basis: str = "wachspress", metric=None, denominator="original"):
    """kind: 'hex82' (config 0/1 mesh), 'hex20', 'quad40', 'ico3', 'ico5' (QU240), 'ico7' (QU60)."""
    if kind.startswith("hex"):
        n = int(kind[3:])
        ny = {82: 94}.get(n, n + 2)
        mesh = meshgen.planar_hex(n, ny, 16000.0)
    elif kind.startswith("quad"):
        n = int(kind[4:])
        mesh = meshgen.planar_quad(n, n, 16000.0)
    elif kind.startswith("ico"):
        mesh = meshgen.icosphere(int(kind[3:]))
    else:
        raise ValueError(kind)
    var = oracle.init_variational(mesh, basis=basis, metric=metric, denominator=denominator)
    return mesh, var


def step_case(mesh, state_kind="auto", config_dt=3600.0, **kw):
    if state_kind == "auto":
        state_kind = "A" if mesh.on_a_sphere else "square"
    if state_kind == "square":
        m2 = mesh
        state = synthetic.square_state(m2)
        # the square forcing is written for Lx = 1.28e6; rescale positions of smaller test meshes
        if abs(mesh.Lx - 1.28e6) > 1.0:
            scaled = copy.copy(mesh)
            scaled = meshgen.Mesh(mesh)
            scaled.xCell = mesh.xCell * (1.28e6 / mesh.Lx)
            scaled.yCell = mesh.yCell * (1.28e6 / mesh.Ly)
            state = synthetic.square_state(scaled)
    else:
        state = synthetic.sphere_state(mesh, kind=state_kind)
    return synthetic.pre_subcycle(mesh, state, config_dt, **kw)


def clone_step(step):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in step.items()}


def run_oracle(mesh, var, step, opts, n_sub):
    s = clone_step(step)
    oracle.subcycle_velocity_solver(mesh, var, s, opts, n_sub)
    return s


def run_device(mesh, var, step, opts, n_sub, **kw):
    from mpas_seaice_b200 import host
    solver = host.EvpSolver(mesh, var, opts, **kw)
    try:
        solver.update_step(step)
        solver.run_subcycles(n_sub)
        out = solver.fetch()
        ms = solver.last_run_ms()
    finally:
        solver.destroy()
    out["_ms"] = ms
    return out


def rel_max_err(a, b, mask=None):
    """max |a-b| / max |b| over the masked entries (the north star's relative max-norm)."""
    if mask is not None:
        a = a[mask]
        b = b[mask]
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / scale)


def masks_for(mesh, step):
    nC, nV = mesh.nCells, mesh.nVertices
    cm = np.zeros(nC + 1, dtype=bool)
    cm[:nC] = step["solveStress"][:nC] == 1
    slot = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:, None]
    vm = np.zeros(nV + 1, dtype=bool)
    vm[:nV] = step["solveVelocity"][:nV] == 1
    return cm[:, None] & slot, vm


# ---------------------------------------------------------------------------------------------
# decomposed runs (host-side emulation of the MPI / NCCL path with the oracle as the block solver)
# ---------------------------------------------------------------------------------------------

def make_blocks(mesh, n_parts, method="auto", n_halos=None, basis="wachspress"):
    """[(block mesh, block var)] + the evp_set_halo lists of every rank."""
    from mpas_seaice_b200 import partition
    part = partition.partition_cells(mesh, n_parts, method)
    blocks = [partition.build_block(mesh, part, r, n_halos) for r in range(n_parts)]
    requests = {r: partition.halo_requests(b) for r, b in enumerate(blocks)}
    lists = [partition.exchange_lists(b, requests) for b in blocks]
    return part, blocks, lists


def exchange_halos(fields_by_rank, lists, names=("uVelocity", "vVelocity")):
    """What the per-subcycle halo exchange does (velocity_solver.F:2543-2584): owner -> halo copies."""
    n = len(lists)
    for r in range(n):
        nbr, soff, sidx, roff, ridx = lists[r]
        for k, q in enumerate(nbr):
            q = int(q)
            qn, qsoff, qsidx, qroff, qridx = lists[q]
            kk = int(np.nonzero(qn == r)[0][0])
            src = qsidx[qsoff[kk]:qsoff[kk + 1]] - 1
            dst = ridx[roff[k]:roff[k + 1]] - 1
            assert src.shape == dst.shape
            for name in names:
                fields_by_rank[r][name][dst] = fields_by_rank[q][name][src]


def run_oracle_blocks(mesh, step, opts, n_sub, n_parts, method="auto", n_halos=None, basis="wachspress"):
    """n_sub subcycles on n_parts blocks, halo exchange after every subcycle; returns the gathered
    global fields (owned entries of every block) and the blocks."""
    from mpas_seaice_b200 import partition
    part, blocks, lists = make_blocks(mesh, n_parts, method, n_halos)
    nC, nV = mesh.nCells, mesh.nVertices
    bvars = [oracle.init_variational(b, basis=basis) for b in blocks]
    bsteps = [partition.restrict_step(b, step, nC, nV) for b in blocks]
    bopts = [dict(opts, nVerticesSolve=int(b.nVerticesSolve)) for b in blocks]
    for _ in range(n_sub):
        for b, v, s, o in zip(blocks, bvars, bsteps, bopts):
            oracle.subcycle_velocity_solver(b, v, s, o, 1)
        exchange_halos(bsteps, lists)
    out = {}
    for k in COMPARE_CELL:
        out[k] = np.zeros_like(step[k])
    for k in COMPARE_VERTEX:
        out[k] = np.zeros_like(step[k])
    for b, s in zip(blocks, bsteps):
        for k in COMPARE_CELL:
            partition.scatter_owned(b, s[k], out[k], "cell")
        for k in COMPARE_VERTEX:
            partition.scatter_owned(b, s[k], out[k], "vertex")
    return out, blocks, lists
