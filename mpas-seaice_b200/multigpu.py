"""One block per rank (= per GPU): what a decomposed MPAS-Seaice host hands to the C-ABI on each rank.

The reference reaches this state through the MPAS framework (block creator reading
graph.info.part.N, dmpar exchange lists) and the halo exchanges of the pre-subcycle
(src/shared/mpas_seaice_velocity_solver.F:842-861, 919-937, 1298-1317, 1480-1498, 1597-1616,
2087-2106); here every rank generates the same global synthetic mesh and per-step fields
deterministically, keeps only its block (partition.build_block / restrict_step) and frees the rest.
The only inter-rank traffic at set-up is the halo request lists and the 128-byte NCCL id, both
through the host's own communicator (torch.distributed here, MPI in the Fortran host).
"""
from __future__ import annotations

import time

import numpy as np

from . import partition, variational_init, workloads


def bind_to_gpu_numa(device_index: int, verbose=None):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (what `mpirun --bind-to` / `numactl` do for the
    Fortran host).  Pages are placed by first touch, so the host arrays allocated afterwards -- the ones the per-step
    copies read and write -- live in the memory next to the GPU's PCIe root instead of crossing the socket link.
    Silently does nothing when the topology cannot be read (no NVML, no sysfs)."""
    import os
    try:
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[0].isdigit() else device_index
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        except Exception:  # noqa: BLE001
            import subprocess
            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                                 capture_output=True, text=True, timeout=20, check=True).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            if verbose:
                verbose(f"GPU {device_index} ({bus}): bound to {len(cpus)} CPUs of its NUMA node ({spec})")
        return sorted(cpus)
    except Exception as e:  # noqa: BLE001 -- an optimisation only
        if verbose:
            verbose(f"NUMA binding skipped: {type(e).__name__}: {e}")
        return None


def build_rank_workload(name, rank, world, dist=None, verbose=None, method="auto", n_halos=None, state="A"):
    """Like workloads.build() but for the block of ``rank`` out of ``world``."""
    log = verbose or (lambda *a: None)
    w = _build_global(name, state, rank, world, dist, verbose)
    gmesh, gstep = w["mesh"], w["step"]
    nCg, nVg = int(gmesh.nCells), int(gmesh.nVertices)
    t0 = time.time()
    part = partition.partition_cells(gmesh, world, method)
    blk = partition.build_block(gmesh, part, rank, n_halos)
    step = partition.restrict_step(blk, gstep, nCg, nVg)
    cells = partition.restrict_step(blk, w["cells"], nCg, nVg)
    cells["iceAreaCell"] = cells["iceAreaCellInitial"]          # one array on the host, uploaded once
    interior = partition.restrict_field(blk, w["interiorVertex"], nCg, nVg)    # the global flags, halo included
    nVs, nCs = int(blk.nVerticesSolve), int(blk.nCellsSolve)
    active = (int((step["solveStress"][:nCs] == 1).sum()), int((step["solveVelocity"][:nVs] == 1).sum()))
    active_local_cells = int((step["solveStress"][:blk.nCells] == 1).sum())
    del gstep, w["step"], w["mesh"], gmesh, w["cells"], w["interiorVertex"]
    static = variational_init.init_static(blk)
    requests = partition.halo_requests(blk)
    log(f"partition '{method}' into {world}: block {rank} has {nCs} owned + {blk.nCells - nCs} halo cells, "
        f"{nVs} owned + {blk.nVertices - nVs} halo vertices ({time.time() - t0:.1f}s)")
    return dict(name=name, mesh=blk, static=static, step=step, opts=w["opts"], config_dt=w["config_dt"], cells=cells,
                interiorVertex=interior,
                nVerticesSolve=nVs, nCellsSolve=nCs, active=active, active_local_cells=active_local_cells,
                global_cells=nCg, global_vertices=nVg, requests=requests,
                partition=f"{method} cell-graph partition into {world} blocks, "
                          f"{blk.nHalos} halo layer(s), vertex owner = first cell of cellsOnVertex")


def _build_global(name, state, rank, world, dist, verbose):
    """The global synthetic workload on every rank.  Large spheres are generated ONCE per machine: rank 0 fills
    the mesh cache (workloads.mesh_cache_path) if needed, the others read it after a barrier -- a 10 M-cell mesh
    takes ~1 min of numpy per generation, eight copies of that work on one host would dominate the set-up."""
    import os
    level = workloads.SPHERES.get(name, (0, 0))[0]
    path = workloads.mesh_cache_path(level)
    if dist is None or world == 1 or path is None:
        return workloads.build(name, state=state, verbose=verbose, with_static=False)
    w = None
    if rank == 0 or os.path.exists(path):
        w = workloads.build(name, state=state, verbose=verbose, with_static=False)
    dist.barrier()
    if w is None:
        w = workloads.build(name, state=state, verbose=verbose, with_static=False)
    return w


def gather_requests(requests, rank, world, dist):
    """All ranks' halo request lists, through the host communicator."""
    if dist is None or world == 1:
        return {rank: requests}
    allreq = [None] * world
    dist.all_gather_object(allreq, requests)
    return {q: r for q, r in enumerate(allreq)}


def attach_halo(solver, w, rank, world, dist):
    """evp_comm_init + evp_set_halo on this rank's handle."""
    ids = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    solver.comm_init(rank, world, ids[0])
    lists = partition.exchange_lists(w["mesh"], gather_requests(w["requests"], rank, world, dist))
    solver.set_halo(*lists)
    return lists
