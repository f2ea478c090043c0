"""One rank of a decomposed run (launched by test_multirank.py with RANK / WORLD_SIZE / MASTER_* set).

  mode 'oracle' : backend gloo, the CPU oracle solves the block, halo exchange by dist.isend/irecv
                  following the very evp_set_halo lists -- covers the N>1 host logic without a GPU;
  mode 'gpu'    : backend nccl, the block is solved by libevp_b200.so; 'gpu-p2p' insists on the peer-to-peer halo
                  exchange fused into the vertex kernel (EVP_B200_HALO=p2p: evp_set_halo fails if it cannot be
                  set up), 'gpu-nccl' on the in-graph ncclSend/ncclRecv exchange, plain 'gpu' takes what
                  evp_set_halo chooses.
Rank 0 gathers the owned results and writes them to an .npz for the parent test to compare."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage(msg):
    print(f"[rank {os.environ.get('RANK')}] {msg}", flush=True)


def main():
    mode, name, nsub, out_path = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    after = int(os.environ.get("EVP_RANK_TRACE_AFTER", "0"))
    if after > 0:
        import faulthandler
        faulthandler.dump_traceback_later(after, exit=True)
    method = sys.argv[5] if len(sys.argv) > 5 else "auto"
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    from mpas_seaice_b200 import multigpu, partition
    import common
    full = mode == "gpu-full"            # pre-subcycle + subcycle + post-subcycle on the device (state B, 2 halo layers)
    if mode in ("gpu-p2p", "gpu-nccl"):
        os.environ["EVP_B200_HALO"] = mode[4:]
    if mode.startswith("gpu"):
        mode = "gpu"
    if mode == "gpu":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo")
    w = multigpu.build_rank_workload(name, rank, world, dist, method=method, n_halos=2 if full else None,
                                     state="B" if full else "A")
    blk, step, opts = w["mesh"], w["step"], w["opts"]
    if mode == "gpu":
        from mpas_seaice_b200 import host
        solver = host.EvpSolver(blk, w["static"], opts, device=rank,
                                local_coords=(w["static"]["xLocal"], w["static"]["yLocal"]),
                                n_vertices_solve=w["nVerticesSolve"], n_cells_solve=w["nCellsSolve"])
        stage("evp_create done")
        multigpu.attach_halo(solver, w, rank, world, dist)
        stage("evp_comm_init + evp_set_halo done: halo exchange " + solver.halo_mode())
        post = {}
        if full:
            solver.set_mesh_ext(blk, w["interiorVertex"])
            solver.pre_subcycle(w["cells"], cold_start=True)
            solver.run_subcycles(nsub)
            post = solver.post_subcycle(names=common.POST_CELL + common.POST_VERTEX)
        else:
            solver.update_step(step)
            solver.run_subcycles(nsub)
        stage("evp_run_subcycles enqueued")
        res = solver.fetch()
        res.update(post)
        stage("evp_fetch done")
        solver.destroy()
    else:
        import oracle
        var = oracle.init_variational(blk)
        lists = partition.exchange_lists(blk, multigpu.gather_requests(w["requests"], rank, world, dist))
        nbr, soff, sidx, roff, ridx = lists
        o = dict(opts, nVerticesSolve=w["nVerticesSolve"])
        for _ in range(nsub):
            oracle.subcycle_velocity_solver(blk, var, step, o, 1)
            reqs, bufs = [], []
            for k, q in enumerate(nbr):
                s = sidx[soff[k]:soff[k + 1]] - 1
                sb = torch.from_numpy(np.stack([step["uVelocity"][s], step["vVelocity"][s]], axis=1).copy())
                rb = torch.empty((int(roff[k + 1] - roff[k]), 2), dtype=torch.float64)
                reqs.append(dist.isend(sb, int(q)))
                reqs.append(dist.irecv(rb, int(q)))
                bufs.append((k, sb, rb))
            for r in reqs:
                r.wait()
            for k, _, rb in bufs:
                d = ridx[roff[k]:roff[k + 1]] - 1
                step["uVelocity"][d] = rb[:, 0].numpy()
                step["vVelocity"][d] = rb[:, 1].numpy()
        res = step
    nCs, nVs = w["nCellsSolve"], w["nVerticesSolve"]
    payload = dict(cell_id=blk.indexToCellID[:nCs], vertex_id=blk.indexToVertexID[:nVs])
    cell_keys = common.COMPARE_CELL + (common.POST_CELL if full else ())
    vertex_keys = common.COMPARE_VERTEX + (common.POST_VERTEX if full else ())
    for k in cell_keys:
        payload[k] = res[k][:nCs]
    for k in vertex_keys:
        payload[k] = res[k][:nVs]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank == 0:
        nC, nV, M = w["global_cells"], w["global_vertices"], blk.maxEdges
        out = {k: (np.zeros((nC + 1, M)) if k in common.COMPARE_CELL else np.zeros(nC + 1)) for k in cell_keys}
        out.update({k: np.zeros(nV + 1) for k in vertex_keys})
        seen_c, seen_v = np.zeros(nC, dtype=int), np.zeros(nV, dtype=int)
        for p in gathered:
            ci, vi = p["cell_id"].astype(np.int64) - 1, p["vertex_id"].astype(np.int64) - 1
            seen_c[ci] += 1
            seen_v[vi] += 1
            for k in cell_keys:
                out[k][ci] = p[k]
            for k in vertex_keys:
                out[k][vi] = p[k]
        assert np.all(seen_c == 1) and np.all(seen_v == 1)
        np.savez(out_path, **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
