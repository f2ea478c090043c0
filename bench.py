#!/usr/bin/env python
"""bench.py -- EVP momentum subcycle throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload qu7.5|qu15|qu30|qu60|qu240|square]
    python bench.py --impl reference ...      # the CPU restatement on the host cores, same metric

A "step" is one dynamics step of the hot path: config_elastic_subcycle_number = 120 EVP subcycles
(strain -> stress -> stress divergence -> drag -> 2x2 solve [-> halo exchange]) on synthetic ice
(state A, every cell active) with analytic forcing.

One JSON line on stdout (rank 0):
  value          whole-job EVP subcycles/s with all inputs resident in HBM (CUDA-graph replay)
  e2e            same metric through the C-ABI with HOST buffers: evp_update_step (H2D of the step's
                 fields) + evp_run_subcycles(120) + evp_fetch (D2H of all outputs) per step
  roofline       dominant kernel (fused cell kernel): SURVEY 8(d) algorithmic bytes per cell x active
                 cells / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
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


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the CPU restatement (oracle, OpenMP where the reference has !$omp parallel do)
    on all host cores, on the same mesh/state; each step is a bounded sample of SUB subcycles."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    from mpas_seaice_b200 import workloads
    name = args.workload
    w = workloads.build(name, state=args.state, verbose=log)
    mesh, step, opts = w["mesh"], w["step"], w["opts"]
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    t0 = time.time()
    var = oracle.init_variational(mesh)
    log(f"oracle precompute {time.time() - t0:.1f}s")
    sub = args.ref_subcycles
    for _ in range(args.warmup):
        oracle.subcycle_velocity_solver(mesh, var, step, opts, sub)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.subcycle_velocity_solver(mesh, var, step, opts, sub)
    dt = time.perf_counter() - t0
    nC_act, nV_act = workloads.active_counts(w)
    value = sub * args.steps / dt
    line = {
        "impl": "reference", "metric": "evp_subcycles_per_sec", "value": value, "unit": "subcycles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps * (N_ELASTIC / sub), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "vertex_updates_per_sec": value * nV_act,
        "config": {"workload": name, "cells": int(mesh.nCells), "vertices": int(mesh.nVertices),
                   "active_cells": nC_act, "active_vertices": nV_act, "subcycles_per_step": N_ELASTIC, "state": args.state,
                   "basis": "wachspress/dunavant-8", "partition": "none (OpenMP threads share one block)"},
        "cpu_baseline": {"value": value, "unit": "subcycles/s", "cores": cores, "kind": "port",
                         "sample": f"{sub} of {N_ELASTIC} subcycles per step on the full {name} mesh"},
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
    ap.add_argument("--state", default="A", choices=["A", "B"],
                    help="synthetic ice state: A = full cover (every cell active, the roofline case), B = polar caps "
                         "(lat > 70N or < 60S, about 10 %% of the cells: what the masks skip)")
    ap.add_argument("--ref-subcycles", type=int, default=2, help="subcycles per step of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    from mpas_seaice_b200 import host, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
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

    name = args.workload
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
    if world > 1:
        multigpu.attach_halo(solver, w, rank, world, dist)
    log(f"rank {rank}: evp_create + device Wachspress precompute {time.time() - t0:.1f}s, "
        f"{solver.device_bytes() / 2**30:.1f} GiB on device")
    solver.update_step(step)

    nC_act, nV_act = w.get("active") or workloads.active_counts(w)
    if dist is not None:
        t = torch.tensor([nC_act, nV_act], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nC_tot, nV_tot = int(t[0]), int(t[1])
    else:
        nC_tot, nV_tot = nC_act, nV_act

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
    if dist is not None:
        t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms = float(t[0]), float(t[1])
    value = N_ELASTIC * args.steps / wall
    launches = args.steps * solver.launch_count(N_ELASTIC)

    # ---- dominant kernel: live CUDA-event timing of the cell pass ---------------------------------
    cell_ms, vertex_ms, other_ms = solver.profile_passes(20)
    peak, peak_src = measured_peak()
    algo_cell = workloads.ALGO_BYTES_PER_CELL * nC_act
    achieved = algo_cell / (cell_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("workload") == name and world == 1:
                traffic = tj.get("cell_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "evp_cell_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_cell, "kernel_ms": cell_ms,
                "vertex_kernel_ms": vertex_ms, "other_ms": other_ms,
                "subcycle_frac_of_peak": (workloads.ALGO_BYTES_PER_CELL_SUBCYCLE * nC_act /
                                          ((cell_ms + vertex_ms + other_ms) * 1e-3) / 1e9) / peak}

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
        n_e2e = max(2, min(args.steps, 3))

        def timed(fn):
            fn()                                      # first call allocates / page-locks the host arrays
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                fn()
            barrier()
            wall_ = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([wall_], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                wall_ = float(t[0])
            return wall_

        out = {}

        def narrow_step():
            solver.update_step(step)
            solver.run_subcycles(N_ELASTIC)
            solver.fetch(into=out)

        w_narrow = timed(narrow_step)
        e2e_narrow = {"value": N_ELASTIC * n_e2e / w_narrow, "unit": "subcycles/s",
                      "h2d_bytes_per_step": step_bytes(step, host.STEP_FIELDS),
                      "d2h_bytes_per_step": int(sum(a.nbytes for a in out.values())),
                      "ms_per_step": 1e3 * w_narrow / n_e2e}
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

        w_wide = timed(wide_step)
        uniq = {id(a): a.nbytes for a in cells.values() if a is not None}
        e2e = {"value": N_ELASTIC * n_e2e / w_wide, "unit": "subcycles/s",
               "h2d_bytes_per_step": int(sum(uniq.values())),
               "d2h_bytes_per_step": int(sum(a.nbytes for a in post.values())),
               "steps": n_e2e, "ms_per_step": 1e3 * w_wide / n_e2e,
               "call": "evp_pre_subcycle(cell fields) + evp_run_subcycles(120) + evp_post_subcycle(u, v, divergence, "
                       "shear, ridgeConvergence, ridgeShear, oceanStressCellU/V)",
               "host_memory": "page-locked via cudaHostRegister (EVP_FLAG_PIN_HOST)"}
        solver.update_step(step)                    # back to the benchmark state for the legs below

    # ---- CPU baseline on the same mesh and state (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            cores = os.cpu_count() or 1
            oracle.set_num_threads(cores)
            t0 = time.time()
            var = dict(static)
            var.update(solver.fetch_basis())         # the basis the GPU uses (bit-identical to the oracle's)
            log(f"fetch_basis for the CPU baseline {time.time() - t0:.1f}s")
            sub = args.ref_subcycles
            cstep = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in step.items()}
            oracle.subcycle_velocity_solver(mesh, var, cstep, opts, 1)      # warm-up
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                oracle.subcycle_velocity_solver(mesh, var, cstep, opts, sub)
            dt = time.perf_counter() - t0
            cpu = {"value": sub * reps / dt, "unit": "subcycles/s", "cores": cores, "kind": "port",
                   "sample": f"{reps} x {sub} subcycles of the same {name} mesh and state (oracle, OpenMP over "
                             f"cells/vertices as in the reference)"}
            del var
        except Exception as e:  # the oracle is optional test infrastructure
            cpu = {"value": None, "unit": "subcycles/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    solver.destroy()
    if rank == 0:
        line = {
            "metric": "evp_subcycles_per_sec", "value": value, "unit": "subcycles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "vertex_updates_per_sec": value * nV_tot,
            "device_ms_per_step": dev_ms / args.steps,
            "config": {"workload": name, "cells": int(w.get("global_cells", mesh.nCells)),
                       "vertices": int(w.get("global_vertices", mesh.nVertices)),
                       "active_cells": nC_tot, "active_vertices": nV_tot,
                       "subcycles_per_step": N_ELASTIC, "state": args.state, "basis": "wachspress/dunavant-8",
                       "l2": "inputs larger than L2 (no flush needed)" if nC_tot * 2240 > 4 * 126e6
                             else "working set fits L2: flush not applied, see DESIGN.md",
                       "partition": w.get("partition", "none")},
            "clocks": clocks, "e2e": e2e, "e2e_subcycle_boundary": e2e_narrow, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
