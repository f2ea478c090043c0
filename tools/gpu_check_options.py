"""Seconds-long device check of the transport options (ir_set_checks, ir_normal_vectors, ir_run_upwind) against the
oracle -- the same assertions as the `cuda` legs of tests/test_transport_options.py, without pytest or torch, for a
GPU box with little time.  Writes gpurun_out/options_check.json and exits non-zero on the first failure."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

t0 = time.time()
import mpas_seaice_b200  # noqa: E402,F401
from mpas_seaice_b200 import ir_host, variational_init  # noqa: E402
from oracle import ir, upwind  # noqa: E402
from test_oracle_ir import case, smooth_divergent_velocity, _random_state  # noqa: E402
from test_ir_parity import clone  # noqa: E402
import test_transport_options as T  # noqa: E402

out = {"steps": []}
LIB = ir_host.LIB_PATH


def note(name, **kw):
    kw["name"] = name
    kw["t"] = round(time.time() - t0, 2)
    out["steps"].append(kw)
    print(json.dumps(kw), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "options_check.json"), "w"), indent=1)


# 1. normal vectors: device trig against the host libm
for kind in ("hex12", "ico3", "band48"):
    mesh, irf, _ = case(kind)
    iv = variational_init.interior_vertex(mesh)
    for rm in (True, False):
        ref = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=rm)
        got = ir_host.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=rm, lib_path=LIB)
        errs = {}
        for key in ref:
            d = np.abs(got[key] - ref[key])
            if key.startswith("lat"):
                errs[key] = float(d.max())
            else:
                errs[key + "_1"] = float(d[..., 0].max())
                errs[key + "_2"] = float(d[..., 1].max())
                errs[key + "_2w"] = float((d[..., 1] * np.maximum(np.abs(ref[key][..., 1]), 1.5e-8)).max())
        note("normals", kind=kind, remove_metric_terms=rm, **errs)
        T._assert_normals_close(got, ref, exact=False)

# 2. checks
for kind in ("hex16", "quad16"):
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    tracers = _random_state(mesh, np.random.default_rng(21))
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, dev = clone(tracers), clone(tracers)
    s = ir_host.IrTransport(mesh, irf, geom, tracers[0].array.shape[1], lib_path=LIB)
    s.set_tracers(dev)
    s.set_checks(1, 1)
    for step in range(3):
        d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, check=False, conservation_check=1, monotonicity_check=2)
        rc = s.run(dev, u, v, 3600.0, check=False)
        rep = s.check_report()
        same = all(np.array_equal(x.array[:nC], y.array[:nC]) for x, y in zip(ref, dev))
        mono_same = [rep["monotonicityViolated"], rep["monoTracer"], rep["monoLayer"], rep["monoCategory"], rep["monoCell"]] == \
            [int(x) for x in d_ref["monoErr"]]
        note("checks", kind=kind, step=step, oracle_rc=int(d_ref["error"]), rc=int(rc), fields_identical=bool(same), mono_same=bool(mono_same))
        assert same and mono_same and rc == {0: 0, 10: 15, 9: 14}[int(d_ref["error"])]
        T._assert_sums_close(d_ref, s, dev)
    s.destroy()

# 3. upwind
for kind in ("quad16", "ico3"):
    mesh, irf, geom, interior, nve = T._upwind_setup(kind)
    nCS = (2 * mesh.nCells) // 3
    var = T._upwind_state(mesh, np.random.default_rng(11))
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, dev = T._clone_vars(var), T._clone_vars(var)
    s = ir_host.IrTransport(mesh, irf, geom, var[0].array.shape[1], n_cells_solve=nCS, lib_path=LIB)
    s.set_upwind_mesh(interior, mesh.dvEdge, nve)
    for step in range(3):
        d = upwind.run(mesh, irf["verticesOnEdge"], interior, nve, ref, u, v, 3600.0, n_cells_solve=nCS, diagnostics=True)
        s.run_upwind(dev, u, v, 3600.0)
        same = all(np.array_equal(x.array, y.array) for x, y in zip(ref, dev))
        fl = all(np.array_equal(s.upwind_fluxes(i)[0][:mesh.nEdges], d["edgeFlux"][i][:mesh.nEdges]) for i in range(len(var)))
        note("upwind", kind=kind, step=step, fields_identical=bool(same), fluxes_identical=bool(fl), kernel_ms=s.last_run_ms())
        assert same and fl
    s.destroy()
# 4. the reference-executed transport fixtures (tests/golden/ir, tests/golden/options) on the device
import glob
import test_ir_parity as TP
for path in TP.REFEXEC_IR:
    TP.test_transport_reproduces_the_reference_executed_steps(path, LIB)
    note("refexec_ir", fixture=os.path.basename(path), ok=True)
for path in TP.REFEXEC_IR_INIT:
    TP.test_geometry_reproduces_the_reference_executed_init(path, LIB)
    note("refexec_irinit", fixture=os.path.basename(path), ok=True)
for path in T.NORMAL_FILES:
    T.test_normal_vectors_reproduce_the_reference_executed_arrays(path, ("cuda", LIB))
    note("refexec_normals", fixture=os.path.basename(path), ok=True)
for path in T.UPWIND_FILES:
    T.test_upwind_reproduces_the_reference_executed_steps(path, LIB)
    note("refexec_upwind", fixture=os.path.basename(path), ok=True)
# 5. the remaining cases of tests/test_transport_options.py, called as plain functions with the shipped library
T.test_checks_pass_and_leave_the_fields_alone("ico3", LIB)
T.test_monotonicity_report_matches_oracle("band48", LIB)
T.test_conservation_report_on_a_partial_block(LIB)
T.test_checks_call_order_and_arguments(LIB)
for kind in ("hex16", "band48"):
    T.test_upwind_matches_oracle(kind, "reference", LIB)
T.test_upwind_call_order_and_arguments(LIB)
for kind in ("quad10",):
    for rm in (True, False):
        T.test_normal_vectors_match_oracle(kind, rm, ("cuda", LIB))
note("transport_options_remaining_cases", ok=True)
note("done", ok=True)
