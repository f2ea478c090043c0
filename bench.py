#!/usr/bin/env python
"""bench.py -- EVP momentum subcycle throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload qu7.5|qu15|qu30|qu60|qu240|square|hexNXxNY]
                    [--scaling strong|weak]
    python bench.py --impl reference ...      # the CPU restatement on the host cores, same metric and config

A "step" is one dynamics step of the hot path: config_elastic_subcycle_number = 120 EVP subcycles
(strain -> stress -> stress divergence -> drag -> 2x2 solve [-> halo exchange]) on synthetic ice
(state A, every cell active) with analytic forcing.

One JSON line on stdout (rank 0):
  value          whole-job EVP subcycles/s with all inputs resident in HBM (CUDA-graph replay)
  e2e            same metric through the C-ABI with HOST buffers: one seaice_run_velocity_solver per step =
                 evp_pre_subcycle (H2D of the step's cell fields) + evp_run_subcycles(120) + evp_post_subcycle (D2H)
  roofline       dominant kernel (fused cell kernel): `frac` = SURVEY 8(d) algorithmic bytes per cell x active
                 cells / its CUDA-event duration / peak; `frac_real_bytes` = the same with the DRAM bytes ncu
                 counted for that kernel (profiles/traffic_r02.json) -- the kernel reads band-compressed
                 gradients, so its real traffic is BELOW the 8(d) model and `frac` can exceed 1
  parity         device against the CPU oracle after the SAME subcycles of the SAME workload, outside every
                 timed region (N = 1; relative max-norm per field, the tolerance asserted is bit-equality)
  checksum       64-bit checksum of the OWNED uVelocity, vVelocity, stress11/22/12 after one fixed 120-subcycle
                 step, hashed with the global ids and summed over the ranks: identical for every --gpus N
                 (the reference's bit-for-bit policy across rank counts, parallelism.py:75-85)
  cpu_baseline   the oracle (kind "port": the reference cannot be built) on the box's host
                 cores for a bounded number of subcycles of the SAME mesh and state
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELASTIC = 120
WEAK_NX, WEAK_NY = 1024, 1280          # --scaling weak: cells per GPU of the planar hex series (1 310 720)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.first = 0
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are used only if none arrives afterwards."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[self.first:] or self.lines[-3:]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def step_bytes(step, names):
    return int(sum(step[n].nbytes for n in names if step.get(n) is not None))


def workload_name(args, world):
    """--scaling weak: a planar hex mesh with a fixed number of cells per GPU (the icosahedral family only
    has sizes that differ by a factor of 4)."""
    if args.scaling == "weak":
        return f"hex{WEAK_NX}x{WEAK_NY * world}"
    return args.workload


def config_dict(name, args, world, cells, vertices, active_cells, active_vertices):
    """The `config` object: the SAME keys and strings in both arms (our arm and --impl reference)."""
    if world > 1:
        part = (f"cell-graph partition into {world} blocks (contiguous blocks of the space-filling-curve cell order on "
                f"the icosphere, RCB otherwise), 1 halo layer, vertex owner = first cell of cellsOnVertex")
    else:
        part = "none"
    return {"workload": name, "cells": int(cells), "vertices": int(vertices),
            "active_cells": int(active_cells), "active_vertices": int(active_vertices),
            "subcycles_per_step": N_ELASTIC, "state": args.state,
            "basis": "hex planar: wachspress/dunavant-8, no metric terms" if name.startswith(("hex", "square"))
                     else "wachspress/dunavant-8",
            "l2": "inputs larger than L2 (no flush needed)" if active_cells * 2240 / max(world, 1) > 4 * 126e6
                  else "working set fits L2: flush not applied, see DESIGN.md",
            "partition": part, "scaling": args.scaling}


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the CPU restatement (oracle, OpenMP where the reference has !$omp parallel do) on all
    host cores, on the same mesh/state.  Each timed step is a bounded SAMPLE of the 120 subcycles of a dynamics
    step, sized from the warm-up rate so that the timed region stays near --ref-budget seconds; `ms_per_step` is
    the measured time of such a sampled step (so steps x ms_per_step is the real timed region), the time of a full
    120-subcycle step is reported as `ms_per_full_step` (extrapolated unless the sample is the whole step)."""
    if rank != 0:
        return
    import oracle
    from mpas_seaice_b200 import workloads
    name = workload_name(args, world)
    w = workloads.build(name, state=args.state, verbose=log, with_static=False)     # no product library involved
    mesh, step, opts = w["mesh"], w["step"], w["opts"]
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    t0 = time.time()
    var = oracle.init_variational(mesh)
    log(f"oracle precompute {time.time() - t0:.1f}s")
    t0 = time.perf_counter()
    for _ in range(max(args.warmup, 1)):
        oracle.subcycle_velocity_solver(mesh, var, step, opts, 2)
    rate = 2 * max(args.warmup, 1) / (time.perf_counter() - t0)
    sub = args.ref_subcycles or int(max(2, min(N_ELASTIC, args.ref_budget * rate / max(args.steps, 1))))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.subcycle_velocity_solver(mesh, var, step, opts, sub)
    dt = time.perf_counter() - t0
    nC_act, nV_act = workloads.active_counts(w)
    value = sub * args.steps / dt
    line = {
        "impl": "reference", "metric": "evp_subcycles_per_sec", "value": value, "unit": "subcycles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "vertex_updates_per_sec": value * nV_act,
        "subcycles_per_timed_step": sub, "ms_per_full_step": 1e3 * N_ELASTIC / value,
        "extrapolated_from": None if sub == N_ELASTIC else sub,
        "config": config_dict(name, args, world, mesh.nCells, mesh.nVertices, nC_act, nV_act),
        "cpu_baseline": {"value": value, "unit": "subcycles/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} timed steps of {sub} of the {N_ELASTIC} subcycles of a dynamics step on "
                                   f"the full {name} mesh, one block shared by {cores} OpenMP threads "
                                   f"(oracle/evp_oracle.c, gcc -O2 -ffp-contract=off)"},
        "e2e": {"value": value, "unit": "subcycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("EVP_BENCH_WORKLOAD", "qu7.5"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the same --workload mesh on every N (default, the headline); weak: a planar hex mesh of "
                         f"{WEAK_NX * WEAK_NY} cells per GPU")
    ap.add_argument("--state", default="A", choices=["A", "B"],
                    help="synthetic ice state: A = full cover (every cell active, the roofline case), B = polar caps "
                         "(lat > 70N or < 60S, about 10 %% of the cells: what the masks skip)")
    ap.add_argument("--ref-subcycles", type=int, default=0,
                    help="subcycles per timed step of the CPU sample (0 = sized from --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=100.0, help="seconds of timed CPU work of --impl reference")
    ap.add_argument("--cpu-subcycles", type=int, default=2, help="subcycles per repetition of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="also skips the parity leg (it shares the oracle run)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-checksum", action="store_true")
    ap.add_argument("--halo", default=os.environ.get("EVP_B200_HALO", "auto"), choices=["auto", "p2p", "nccl"],
                    help="per-subcycle halo exchange: peer stores over NVLink fused into the vertex kernel, or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    from mpas_seaice_b200 import checksum, host, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1 and not os.environ.get("EVP_B200_NO_NUMA_BIND"):
        from mpas_seaice_b200 import multigpu as _mg
        _mg.bind_to_gpu_numa(local_rank, verbose=log if rank == 0 else None)
    os.environ["EVP_B200_HALO"] = args.halo
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(*vals):
        if dist is None:
            return vals if len(vals) > 1 else vals[0]
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in t) if len(vals) > 1 else float(t[0])

    name = workload_name(args, world)
    if world > 1:
        from mpas_seaice_b200 import multigpu
        w = multigpu.build_rank_workload(name, rank, world, dist, verbose=log if rank == 0 else None, state=args.state)
    else:
        w = workloads.build(name, state=args.state, verbose=log)
    mesh, static, step, opts = w["mesh"], w["static"], w["step"], w["opts"]

    t0 = time.time()
    solver = host.EvpSolver(mesh, static, opts, device=local_rank, pin_host=True,
                            local_coords=(static["xLocal"], static["yLocal"]),
                            n_vertices_solve=w.get("nVerticesSolve"), n_cells_solve=w.get("nCellsSolve"))
    halo_mode = "none"
    if world > 1:
        multigpu.attach_halo(solver, w, rank, world, dist)
        halo_mode = solver.halo_mode()
    log(f"rank {rank}: evp_create + device Wachspress precompute {time.time() - t0:.1f}s, "
        f"{solver.device_bytes() / 2**30:.1f} GiB on device, halo exchange: {halo_mode}")
    solver.update_step(step)

    nC_act, nV_act = w.get("active") or workloads.active_counts(w)
    if dist is not None:
        t = torch.tensor([nC_act, nV_act], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nC_tot, nV_tot = int(t[0]), int(t[1])
    else:
        nC_tot, nV_tot = nC_act, nV_act

    # ---- checksum of the owned results of one fixed dynamics step (outside the timed regions) -------------
    csum = None
    if not args.no_checksum:
        solver.run_subcycles(N_ELASTIC)
        got = solver.fetch(names=("uVelocity", "vVelocity", "stress11", "stress22", "stress12"))
        part = checksum.owned_checksum(mesh, got)
        # the arrays of this fetch were page-locked (EVP_FLAG_PIN_HOST): unregister them BEFORE they are freed -- a
        # stale registration under a recycled address makes a later copy fail or land in the old pages
        solver.release_host_memory()
        del got
        if dist is not None:
            parts = [None] * world
            dist.all_gather_object(parts, part)
        else:
            parts = [part]
        csum = {"value": "%016x" % checksum.combine(parts), "subcycles": N_ELASTIC,
                "fields": "owned uVelocity, vVelocity, stress11, stress22, stress12 hashed with their global ids; "
                          "must be identical for every --gpus N on this workload and state"}
        solver.update_step(step)                    # back to the initial state

    # ---- device-resident throughput -----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()                  # nvidia-smi needs ~0.1-0.2 s to deliver its first sample: start it before the
    for _ in range(args.warmup):     # warm-up so that short timed regions (small meshes, many GPUs) are covered too
        solver.run_subcycles(N_ELASTIC)
    solver.synchronize()
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        solver.run_subcycles(N_ELASTIC)
        solver.synchronize()
        dev_ms += solver.last_run_ms()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    wall, dev_ms = all_max(wall, dev_ms)
    value = N_ELASTIC * args.steps / wall
    launches = args.steps * solver.launch_count(N_ELASTIC)

    # ---- dominant kernel: live CUDA-event timing of the cell pass ---------------------------------
    cell_ms, vertex_ms, other_ms = solver.profile_passes(20)
    cell_ms, vertex_ms, other_ms = all_max(cell_ms, vertex_ms, other_ms)
    peak, peak_src = measured_peak()
    planar = name.startswith(("hex", "square"))
    algo_per_cell = workloads.ALGO_BYTES_PER_CELL_PLANAR if planar else workloads.ALGO_BYTES_PER_CELL
    algo_per_sub = algo_per_cell + 2 * workloads.ALGO_BYTES_PER_VERTEX
    nC_local = w.get("active_local_cells", nC_act)      # a rank's cell pass covers its halo cells too
    algo_cell = algo_per_cell * nC_local
    achieved = algo_cell / (cell_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("workload") == name and world == 1 and tj.get("state", "A") == args.state:
                traffic = tj.get("cell_kernel_dram_bytes_per_launch")
                traffic_src = tj.get("source")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_model": f"SURVEY 8(d): {algo_per_cell:.0f} B per active cell (dense basisGradientU/V counted)",
                "traffic": traffic,
                "frac_real_bytes": (traffic / (cell_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "frac_real_bytes_model": traffic_src,
                "kernel": "evp_cell_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_cell, "kernel_ms": cell_ms,
                "vertex_kernel_ms": vertex_ms, "other_ms": other_ms,
                "subcycle_frac_of_peak": (algo_per_sub * nC_local /
                                          ((cell_ms + vertex_ms + other_ms) * 1e-3) / 1e9) / peak,
                "graph_ms_per_subcycle": dev_ms / args.steps / N_ELASTIC}

    # ---- end to end through the C-ABI with HOST buffers ---------------------------------------------------
    # e2e = one seaice_run_velocity_solver per step through the widened boundary: evp_pre_subcycle (H2D of the
    # step's CELL fields from page-locked host arrays, pre-subcycle on the device) + evp_run_subcycles(120) +
    # evp_post_subcycle (post-subcycle on the device, D2H of what the model consumes every step: u, v ->
    # advection; divergence, shear, ridgeConvergence, ridgeShear -> ridging; oceanStressCellU/V -> coupler).
    # e2e_subcycle_boundary = the narrower boundary of round-1's first cut: evp_update_step (H2D of all 21
    # vertex/cell step fields) + evp_run_subcycles(120) + evp_fetch (D2H of all 12 outputs).
    e2e = None
    e2e_narrow = None
    if not args.no_e2e:
        def timed(fn, n):
            fn()                                      # first call allocates / page-locks the host arrays
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            barrier()
            return all_max(time.perf_counter() - t0)

        out = {}

        def narrow_step():
            solver.update_step(step)
            solver.run_subcycles(N_ELASTIC)
            solver.fetch(into=out)

        n_narrow = max(2, min(args.steps, 3))
        w_narrow = timed(narrow_step, n_narrow)
        e2e_narrow = {"value": N_ELASTIC * n_narrow / w_narrow, "unit": "subcycles/s",
                      "h2d_bytes_per_step": step_bytes(step, host.STEP_FIELDS),
                      "d2h_bytes_per_step": int(sum(a.nbytes for a in out.values())),
                      "steps": n_narrow, "ms_per_step": 1e3 * w_narrow / n_narrow}
        # `out` stays alive until solver.destroy(): with EVP_FLAG_PIN_HOST its arrays are registered with CUDA for
        # the life of the handle (the contract of the flag); freeing them earlier leaves stale registrations that
        # can collide with later device allocations (seen as 'resource already mapped' at N = 8)

        cells = w["cells"]
        solver.set_mesh_ext(mesh, w["interiorVertex"])
        post = {}
        first = [True]

        def wide_step():
            solver.pre_subcycle(cells, cold_start=first[0])
            first[0] = False
            solver.run_subcycles(N_ELASTIC)
            solver.post_subcycle(into=post)

        n_e2e = max(2, args.steps)
        w_wide = timed(wide_step, n_e2e)
        uniq = {id(a): a.nbytes for a in cells.values() if a is not None}
        e2e = {"value": N_ELASTIC * n_e2e / w_wide, "unit": "subcycles/s",
               "h2d_bytes_per_step": int(sum(uniq.values())),
               "d2h_bytes_per_step": int(sum(a.nbytes for a in post.values())),
               "steps": n_e2e, "ms_per_step": 1e3 * w_wide / n_e2e,
               "copy_and_prepost_ms_per_step": 1e3 * w_wide / n_e2e - dev_ms / args.steps,
               "call": "evp_pre_subcycle(cell fields) + evp_run_subcycles(120) + evp_post_subcycle(u, v, divergence, "
                       "shear, ridgeConvergence, ridgeShear, oceanStressCellU/V)",
               "host_memory": "page-locked via cudaHostRegister (EVP_FLAG_PIN_HOST)"}
        solver.update_step(step)                    # back to the benchmark state for the legs below

    # ---- parity against the oracle + CPU baseline on the same mesh and state (rank 0, N = 1 only) ----------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            cores = os.cpu_count() or 1
            oracle.set_num_threads(cores)
            t0 = time.time()
            var = dict(static)
            var.update(solver.fetch_basis())         # the basis the GPU uses (bit-identical to the oracle's)
            log(f"fetch_basis for the CPU baseline {time.time() - t0:.1f}s")
            sub = args.cpu_subcycles
            cstep = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in step.items()}
            # parity: `sub` subcycles from the initial state on both sides (also the oracle's warm-up)
            oracle.subcycle_velocity_solver(mesh, var, cstep, opts, sub)
            solver.run_subcycles(sub)
            names = ("uVelocity", "vVelocity", "stress11", "stress22", "stress12")
            got = solver.fetch(names=names)
            nC, nV = int(mesh.nCells), int(mesh.nVertices)
            vmask = step["solveVelocity"][:nV] == 1
            cmask = (step["solveStress"][:nC] == 1)[:, None] & \
                (np.arange(int(mesh.maxEdges))[None, :] < mesh.nEdgesOnCell[:nC, None])
            errs, exact = {}, True
            for n in names:
                a, b = (got[n][:nV][vmask], cstep[n][:nV][vmask]) if n.endswith("Velocity") else \
                    (got[n][:nC][cmask], cstep[n][:nC][cmask])
                scale = float(np.max(np.abs(b))) if b.size else 0.0
                errs[n] = float(np.max(np.abs(a - b)) / scale) if scale > 0 else float(np.max(np.abs(a), initial=0.0))
                exact = exact and bool(np.array_equal(a, b))
            parity = {"config": name, "state": args.state, "subcycles": sub, "max_rel_err": max(errs.values()),
                      "per_field": errs, "bit_exact": exact, "compared": "device vs oracle/evp_oracle.c, solved "
                      "vertices and every stress point of the solved cells", "tolerance": 1e-10}
            solver.release_host_memory()             # `got` was page-locked by the fetch: unregister before freeing
            del got
            solver.update_step(step)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                oracle.subcycle_velocity_solver(mesh, var, cstep, opts, sub)
            dt = time.perf_counter() - t0
            cpu = {"value": sub * reps / dt, "unit": "subcycles/s", "cores": cores, "kind": "port",
                   "sample": f"{reps} x {sub} subcycles of the same {name} mesh and state (oracle, OpenMP over "
                             f"cells/vertices as in the reference)",
                   "build": "C2 of BASELINE.md section 3: gcc -O2 -ffp-contract=off, all host cores"}
            # BASELINE.md section 3 variants, reported next to it: C1 = one thread ("single CPU rank"), and the
            # -O3 -march=native build (compiled on this machine, contraction at the compiler's default)
            variants = {}
            try:
                oracle.set_num_threads(1)
                t0 = time.perf_counter()
                oracle.subcycle_velocity_solver(mesh, var, cstep, opts, 1)
                variants["C1_single_thread"] = {"value": 1.0 / (time.perf_counter() - t0), "unit": "subcycles/s", "cores": 1,
                                                "sample": "1 subcycle", "build": "gcc -O2 -ffp-contract=off"}
                oracle.set_num_threads(cores)
                fast = oracle.load_variant("o3native")
                oracle.subcycle_velocity_solver(mesh, var, cstep, opts, 1, library=fast)
                t0 = time.perf_counter()
                for _ in range(reps):
                    oracle.subcycle_velocity_solver(mesh, var, cstep, opts, sub, library=fast)
                variants["C2_O3_native"] = {"value": sub * reps / (time.perf_counter() - t0), "unit": "subcycles/s",
                                            "cores": cores, "sample": f"{reps} x {sub} subcycles",
                                            "build": "gcc -O3 -march=native (FMA contraction allowed: timing only)"}
            except Exception as e:  # noqa: BLE001
                variants["error"] = str(e)
            cpu["variants"] = variants
            del var
        except Exception as e:  # the oracle is optional test infrastructure
            cpu = {"value": None, "unit": "subcycles/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    solver.destroy()
    if rank == 0:
        g_cells, g_verts = int(w.get("global_cells", mesh.nCells)), int(w.get("global_vertices", mesh.nVertices))
        line = {
            "metric": "evp_subcycles_per_sec", "value": value, "unit": "subcycles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "vertex_updates_per_sec": value * nV_tot,
            "cell_updates_per_sec_per_gpu": value * nC_tot / world,
            "device_ms_per_step": dev_ms / args.steps,
            "config": config_dict(name, args, world, g_cells, g_verts, nC_tot, nV_tot),
            "halo_exchange": halo_mode,
            "clocks": clocks, "e2e": e2e, "e2e_subcycle_boundary": e2e_narrow, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "checksum": csum,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException as exc:  # noqa: BLE001
        if isinstance(exc, SystemExit) and exc.code in (0, None):
            raise
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        # no interpreter shutdown: the NCCL destructors of a rank that failed would wait for the collectives its
        # peers are blocked in; exiting at once lets the launcher stop the other ranks
        os._exit(1 if not isinstance(exc, SystemExit) else (exc.code if isinstance(exc.code, int) else 1))
