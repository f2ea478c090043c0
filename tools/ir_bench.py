"""Timing of the incremental-remap transport step (include/ir_b200.h) -- a measurement aid for the IR row
(DESIGN.md section 7c), NOT the north-star bench (that is bench.py: EVP subcycles per second).

    python tools/ir_bench.py [--level 7] [--steps 5] [--warmup 2] [--cpu]

Product path only: meshgen -> irmesh (host arrays a Python host lacks) -> ir_init_geometry (device) -> ir_run.
Workload: the reference's standard tracer set (iceAreaCategory, iceVolumeCategory, snowVolumeCategory,
surfaceTemperature, iceEnthalpy, iceSalinity, snowEnthalpy; 5 categories, 7 ice layers, 5 snow layers =
115 (category, layer) rows), smooth ice cover with open water around the equator (where the rotated grid has its
poles), smooth divergent velocity at 30 % of the CFL limit.  Prints one JSON line in bench.py's vocabulary:
  value / kernel_ms_per_step   device time of the five kernels of a step (CUDA events inside ir_run), cell-row updates / s
  e2e                          the same through the C ABI with host buffers (uploads and downloads inside the timed region)
  roofline                     HBM: ALGORITHMIC bytes of a step (per cell-row val in 8 + centre / gradients out
                               and in 48 + edge fluxes of 3 edges out and in 48 + new value out 8 = 112 B; per edge the
                               departure-triangle data out and in 2 x 4 triangles x 112 B) / kernel time / measured peak
  cpu_baseline                 with --cpu-baseline: the oracle (oracle/ir_oracle.c, kind "port") on the same state, few steps
``--cpu`` times the oracle only (the checker timed as a baseline, as bench.py does).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpas_seaice_b200  # noqa: E402,F401
from mpas_seaice_b200 import irmesh, workloads  # noqa: E402


class Tracer:
    def __init__(self, name, array, parent=None, volume_like=False):
        self.name, self.array, self.parent, self.volume_like = name, array, parent, volume_like


def standard_tracers(mesh, n_cat=5, n_ice=7, n_snow=5):
    nC = mesh.nCells
    lat, lon = mesh.latCell[:nC], mesh.lonCell[:nC]
    cover = np.clip((np.abs(lat) - np.deg2rad(20.0)) / np.deg2rad(30.0), 0.0, 1.0) * (0.8 + 0.15 * np.sin(3 * lon))

    def new(nl):
        return np.zeros((nC + 1, n_cat, nl))
    area, ivol, svol, tsfc = new(1), new(1), new(1), new(1)
    ienth, isal, senth = new(n_ice), new(n_ice), new(n_snow)
    for k in range(n_cat):
        a = cover * (0.1 + 0.05 * k) * (1.0 + 0.2 * np.cos(2 * lon + k))
        area[:nC, k, 0] = a
        ivol[:nC, k, 0] = a * (0.5 + 0.6 * k) * (1.0 + 0.1 * np.sin(4 * lat))
        svol[:nC, k, 0] = a * 0.1 * (1.0 + 0.3 * np.cos(lon))
        tsfc[:nC, k, 0] = -10.0 - 5.0 * np.cos(lat) + k
        for l in range(n_ice):
            ienth[:nC, k, l] = -3.0e8 * (1.0 + 0.05 * l + 0.02 * np.sin(lon))
            isal[:nC, k, l] = 3.0 + 0.5 * l + 0.1 * np.cos(lat)
        for l in range(n_snow):
            senth[:nC, k, l] = -1.2e8 * (1.0 + 0.03 * l + 0.02 * np.cos(lon))
    return [Tracer("iceAreaCategory", area), Tracer("iceVolumeCategory", ivol, 0, True),
            Tracer("snowVolumeCategory", svol, 0, True), Tracer("surfaceTemperature", tsfc, 0),
            Tracer("iceEnthalpy", ienth, 1), Tracer("iceSalinity", isal, 1), Tracer("snowEnthalpy", senth, 2)]


def smooth_velocity(mesh, min_edge, dt, cfl=0.3):
    nV = mesh.nVertices
    speed = cfl * min_edge / dt
    lat, lon = mesh.latVertex[:nV], mesh.lonVertex[:nV]
    u, v = np.zeros(nV + 1), np.zeros(nV + 1)
    u[:nV] = speed * np.cos(lat) * np.sin(2 * lon)
    v[:nV] = speed * np.cos(lat) * np.sin(3 * lat) * np.cos(lon)
    return u, v


def upwind_bench(args, mesh, irf, tracers, dt, t_mesh):
    """ir_run_upwind on the same mesh and ice state: 4 variables x 5 categories.  Algorithmic bytes of a step: per
    (variable, category) row the old value in and the new value out per cell (16 B) and the edge flux out per edge (8 B)."""
    from mpas_seaice_b200 import ir_host, variational_init
    from oracle import upwind as oup

    class Var:
        def __init__(self, name, array, parent=None, volume_like=False):
            self.name, self.array, self.parent, self.volume_like, self.child_minimum = name, array, parent, volume_like, 0.0
    variables = [Var(t.name, np.ascontiguousarray(t.array[:, :, 0]), None if t.parent is None else 0, t.volume_like)
                 for t in tracers[:4]]
    geom = ir_host.init_geometry(mesh, irf, rotate=True)
    u, v = smooth_velocity(mesh, geom["minLengthEdgesOnVertex"][:mesh.nVertices].min(), dt)
    iv = variational_init.interior_vertex(mesh)
    nve = ir_host.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    interior = ir_host.interior_edge(mesh)
    nK = variables[0].array.shape[1]
    rows = len(variables) * nK
    rec = dict(metric="upwind_cell_row_updates_per_s", unit="cell-rows/s", cells=mesh.nCells, edges=mesh.nEdges, rows=rows,
               steps=args.steps, warmup=args.warmup, mesh_s=round(t_mesh, 2), data="synthetic", dtype="f64", impl="cuda")
    initial = [x.array.copy() for x in variables]
    solver = ir_host.IrTransport(mesh, irf, geom, nK, rotate=True)
    try:
        solver.set_upwind_mesh(interior, mesh.dvEdge, nve)
        for _ in range(args.warmup):
            solver.run_upwind(variables, u, v, dt)
        dev_ms, t1 = [], time.time()
        for _ in range(args.steps):
            solver.run_upwind(variables, u, v, dt)
            dev_ms.append(solver.last_run_ms())
        wall = (time.time() - t1) / args.steps
        kms = float(np.mean(dev_ms)) or float("nan")      # (0 under the host emulation, which has no clock)
        algo = rows * (16.0 * mesh.nCells + 8.0 * mesh.nEdges)
        peak = 6650.0
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak = float(json.load(open(pk))["hbm_gbs"])
        rec.update(kernel_ms_per_step=round(kms, 4), wall_ms_per_step=round(wall * 1e3, 3), value=mesh.nCells * rows / (kms * 1e-3),
                   higher_is_better=True, gpu_launches=solver.launch_count(),
                   roofline=dict(bound="hbm", achieved=algo / (kms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                 frac=algo / (kms * 1e-3) / 1e9 / peak, algorithmic_bytes_per_step=algo, traffic=None,
                                 kernel="the kernels of one ir_run_upwind (edge velocity, prepare, per variable edge flux + update, finalize)"))
    finally:
        solver.destroy()
    if args.check:
        ref = [oup.Var(x.name, a, x.parent, x.volume_like) for x, a in zip(variables, initial)]
        nve_o = oup.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
        rec["normals_max_abs_diff"] = float(np.abs(nve - nve_o).max())
        for _ in range(args.warmup + args.steps):
            oup.run(mesh, irf["verticesOnEdge"], interior, nve, ref, u, v, dt)
        rec["parity"] = bool(all(np.array_equal(a.array, b.array) for a, b in zip(variables, ref)))
    print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=7, help="icosphere level: 7 = 163 842 cells (QU60), 9 = QU15, 10 = QU7.5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu", action="store_true", help="time the oracle (CPU) instead of the device")
    ap.add_argument("--upwind", action="store_true",
                    help="time config_advection_type = 'upwind' (ir_run_upwind: area, ice / snow volume, surface temperature "
                         "riding on the area) instead of the incremental remapping")
    ap.add_argument("--cpu-baseline", action="store_true", help="add a cpu_baseline object: the oracle timed on 2 steps")
    ap.add_argument("--check", action="store_true",
                    help="after the timed steps, run the oracle on the same initial state and report whether the device "
                         "fields are identical (the oracle as the checker, outside the timed region)")
    args = ap.parse_args()
    dt = 3600.0
    t0 = time.time()
    mesh = workloads._cached_icosphere(args.level) if hasattr(workloads, "_cached_icosphere") else None
    if mesh is None:
        from mpas_seaice_b200 import meshgen
        mesh = meshgen.icosphere(args.level)
    irf = irmesh.ir_fields(mesh)
    t_mesh = time.time() - t0
    tracers = standard_tracers(mesh)
    if args.upwind:
        return upwind_bench(args, mesh, irf, tracers, dt, t_mesh)
    initial = [t.array.copy() for t in tracers] if args.check else None
    n_rows = sum(t.array.shape[1] * t.array.shape[2] for t in tracers)
    rec = dict(metric="ir_cell_row_updates_per_s", unit="cell-rows/s", cells=mesh.nCells, edges=mesh.nEdges, rows=n_rows,
               steps=args.steps, warmup=args.warmup, mesh_s=round(t_mesh, 2), data="synthetic", dtype="f64")
    if args.cpu:
        from oracle import ir
        geom = ir.init_geometry(mesh, irf, rotate=True)          # config_rotate_cartesian_grid: Registry default
        u, v = smooth_velocity(mesh, geom["minLengthEdgesOnVertex"][:mesh.nVertices].min(), dt)
        otr = [ir.Tracer(t.name, t.array, t.parent, t.volume_like) for t in tracers]
        for _ in range(args.warmup):
            ir.run(mesh, irf, geom, otr, u, v, dt, rotate=True)
        t1 = time.time()
        for _ in range(args.steps):
            ir.run(mesh, irf, geom, otr, u, v, dt, rotate=True)
        wall = (time.time() - t1) / args.steps
        rec.update(impl="oracle", cores=os.cpu_count(), wall_ms_per_step=round(wall * 1e3, 3),
                   value=mesh.nCells * n_rows / wall)
    else:
        from mpas_seaice_b200 import ir_host
        t1 = time.time()
        geom = ir_host.init_geometry(mesh, irf, rotate=True)     # config_rotate_cartesian_grid: Registry default
        rec["init_geometry_s"] = round(time.time() - t1, 3)
        u, v = smooth_velocity(mesh, geom["minLengthEdgesOnVertex"][:mesh.nVertices].min(), dt)
        solver = ir_host.IrTransport(mesh, irf, geom, tracers[0].array.shape[1], rotate=True)
        try:
            solver.set_tracers(tracers)
            for _ in range(args.warmup):
                solver.run(tracers, u, v, dt)
            dev_ms, t1 = [], time.time()
            for _ in range(args.steps):
                solver.run(tracers, u, v, dt)
                dev_ms.append(solver.last_run_ms())
            wall = (time.time() - t1) / args.steps
            kms = float(np.mean(dev_ms)) or float("nan")      # (0 under the host emulation, which has no clock)
            algo = mesh.nCells * n_rows * 112.0 + mesh.nEdges * 2 * 4 * 112.0
            peak = 6650.0
            pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(pk):
                peak = float(json.load(open(pk))["hbm_gbs"])
            io_bytes = int(sum(t.array.nbytes for t in tracers))
            rec["kernel_ms"] = {k: round(x, 4) for k, x in solver.last_kernel_ms().items()}     # of the last timed step
            rec.update(impl="cuda", kernel_ms_per_step=round(kms, 4), wall_ms_per_step=round(wall * 1e3, 3),
                       value=mesh.nCells * n_rows / (kms * 1e-3), higher_is_better=True, gpu_launches=solver.launch_count(),
                       e2e=dict(value=mesh.nCells * n_rows / wall, unit="cell-rows/s", h2d_bytes_per_step=io_bytes + 2 * u.nbytes,
                                d2h_bytes_per_step=io_bytes),
                       roofline=dict(bound="hbm", achieved=algo / (kms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                     frac=algo / (kms * 1e-3) / 1e9 / peak, algorithmic_bytes_per_step=algo, traffic=None,
                                     kernel="the five kernels of one ir_run (k_prepare, k_reconstruct_coop, k_triangles, "
                                            "k_fluxes_coop, k_update_coop)"))
        finally:
            solver.destroy()
        if args.check:
            from oracle import ir
            geom_o = ir.init_geometry(mesh, irf, rotate=True)
            otr = [ir.Tracer(t.name, a, t.parent, t.volume_like) for t, a in zip(tracers, initial)]
            for _ in range(args.warmup + args.steps):
                ir.run(mesh, irf, geom_o, otr, u, v, dt, rotate=True)
            nC = mesh.nCells
            rec["parity"] = bool(all(np.array_equal(t.array[:nC], o.array[:nC]) for t, o in zip(tracers, otr)))
            rec["geometry_parity"] = bool(all(np.array_equal(geom[k][:-1], geom_o[k][:-1]) for k in
                                              ("xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge", "remapEdge",
                                               "cellsOnEdgeRemap", "edgesOnEdgeRemap")))
    if args.cpu_baseline and not args.cpu:
        from oracle import ir
        geom_o = ir.init_geometry(mesh, irf, rotate=True)
        otr = [ir.Tracer(t.name, t.array.copy(), t.parent, t.volume_like) for t in tracers]
        ir.run(mesh, irf, geom_o, otr, u, v, dt, rotate=True)
        t1 = time.time()
        for _ in range(2):
            ir.run(mesh, irf, geom_o, otr, u, v, dt, rotate=True)
        cw = (time.time() - t1) / 2
        rec["cpu_baseline"] = dict(value=mesh.nCells * n_rows / cw, unit="cell-rows/s", cores=os.cpu_count(), kind="port",
                                   sample="2 steps of the same mesh and state, oracle/ir_oracle.c (OpenMP)")
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
