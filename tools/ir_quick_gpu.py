"""The shortest possible first contact of the transport kernels with a device: geometry init and two steps of the full
tracer hierarchy on four small meshes, compared bit for bit with the oracle.  No torch, no pytest session: a few
seconds.  Writes gpurun_out/ir_first_device_run.json.

    /usr/local/graft/bin/gpurun --timeout 40 -- 'python tools/ir_quick_gpu.py'
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ir  # noqa: E402  (the checker)
from mpas_seaice_b200 import ir_host  # noqa: E402
import test_oracle_ir as T  # noqa: E402


def main():
    out, t0 = {}, time.time()
    for kind in ("hex16", "quad16", "ico3", "band48"):
        rec = {}
        try:
            mesh, irf, geom = T.case(kind)
            g = ir_host.init_geometry(mesh, irf)
            names = ("xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge", "remapEdge", "cellsOnEdgeRemap",
                     "edgesOnEdgeRemap", "minLengthEdgesOnVertex")
            rec["geometry_identical"] = bool(all(np.array_equal(geom[k][:-1], g[k][:-1]) for k in names)
                                             and all(np.array_equal(geom["geomAvg"][n][:-1], g["geomAvg"][n][:-1])
                                                     for n in ir_host.GEOM_NAMES))
            rng = np.random.default_rng(21)
            ref = T._random_state(mesh, rng)
            dev = [ir.Tracer(t.name, t.array.copy(), t.parent, t.volume_like) for t in ref]
            u, v = T.smooth_divergent_velocity(mesh, geom)
            s = ir_host.IrTransport(mesh, irf, g, 3)
            try:
                s.set_tracers(dev)
                codes, ms = [], []
                for _ in range(2):
                    d = ir.run(mesh, irf, geom, ref, u, v, 3600.0, check=False, diagnostics=True)
                    codes.append((int(d["error"]), int(s.run(dev, u, v, 3600.0, check=False))))
                    ms.append(s.last_run_ms())
                dd = s.diagnostics()
                rec["launches"] = int(s.launch_count())
            finally:
                s.destroy()
            nC = mesh.nCells
            rec["codes"] = codes
            rec["kernel_ms"] = [round(float(x), 4) for x in ms]
            rec["tracers_identical"] = {a.name: bool(np.array_equal(a.array[:nC], b.array[:nC])) for a, b in zip(ref, dev)}
            rec["diagnostics_identical"] = {k: bool(np.array_equal(d[k], dd[k])) for k in dd}
            rec["max_abs_diff"] = {a.name: float(np.abs(a.array[:nC] - b.array[:nC]).max()) for a, b in zip(ref, dev)}
        except Exception as e:  # noqa: BLE001 -- record and go on
            rec["exception"] = "%s: %s" % (type(e).__name__, e)
        out[kind] = rec
        print(kind, json.dumps(rec), flush=True)
    out["seconds"] = round(time.time() - t0, 2)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ir_first_device_run.json"), "w") as f:
        json.dump(out, f, indent=1)
    ok = all(r.get("geometry_identical") and all(r.get("tracers_identical", {"x": False}).values()) for k, r in out.items()
             if isinstance(r, dict))
    print("ALL IDENTICAL" if ok else "DIFFERENCES", out["seconds"], "s")


if __name__ == "__main__":
    main()
