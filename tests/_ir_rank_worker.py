"""One rank of a decomposed incremental-remap run (launched by tests/test_ir_multirank.py with RANK / WORLD_SIZE /
MASTER_* set): builds its block, computes the geometry and runs the transport through the C ABI (argv[1] = path of
the library: the emulation build on CPU, libir_b200.so on a GPU), updates the tracer halos over gloo after every
step, and sends its owned cells to rank 0, which writes the gathered fields to an .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    lib_path, kind, n_steps, out_path = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("IR_RANK_TRACE_AFTER", "150")), exit=True)
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = -1
    if torch.cuda.is_available() and "emu" not in os.path.basename(lib_path):
        device = rank % torch.cuda.device_count()
    dist.init_process_group("gloo")
    from mpas_seaice_b200 import ir_host, partition
    from oracle import ir                      # Tracer containers and the shared synthetic state only
    import test_oracle_ir as T
    mesh, irf, _ = T.case(kind)                # deterministic: the same global mesh on every rank
    part = partition.partition_cells(mesh, world)
    blk = partition.build_block(mesh, part, rank, 2)
    birf = partition.restrict_ir(blk, mesh, irf)
    requests = [None] * world
    dist.all_gather_object(requests, {int(k): v.tolist() for k, v in partition.cell_halo_requests(blk).items()})
    lists = partition.cell_exchange_lists(blk, {q: requests[q] for q in range(world)})
    rng = np.random.default_rng(17)
    tracers = T._random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    geom_global_min = None
    btr = [ir.Tracer(t.name, partition.restrict_field(blk, t.array, mesh.nCells, mesh.nVertices), t.parent, t.volume_like)
           for t in tracers]
    u, v = T.smooth_divergent_velocity(mesh, T.case(kind)[2])
    bu = partition.restrict_field(blk, u, mesh.nCells, mesh.nVertices)
    bv = partition.restrict_field(blk, v, mesh.nCells, mesh.nVertices)
    geom = ir_host.init_geometry(blk, birf, n_cells_solve=blk.nCellsSolve, device=device, lib_path=lib_path)
    solver = ir_host.IrTransport(blk, birf, geom, 2, n_cells_solve=blk.nCellsSolve, device=device, lib_path=lib_path)
    halo = ir_host.TracerHalo(lists, dist)
    try:
        solver.set_tracers(btr)
        for _ in range(n_steps):
            solver.run(btr, bu, bv, 3600.0)
            halo.update(btr)
    finally:
        solver.destroy()
    owned = {t.name: t.array[:blk.nCellsSolve].copy() for t in btr}
    ids = blk.indexToCellID[:blk.nCellsSolve].copy()
    gathered = [None] * world
    dist.gather_object((ids, owned), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        out = {t.name: np.zeros_like(t.array) for t in tracers}
        for gids, vals in gathered:
            for name, a in vals.items():
                out[name][gids.astype(np.int64) - 1] = a
        np.savez(out_path, **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
