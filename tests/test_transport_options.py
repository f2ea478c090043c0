"""The non-default options of the transport row behind include/ir_b200.h, against the oracle:

* config_conservation_check / config_monotonicity_check of the incremental remapping
  (src/shared/mpas_seaice_advection_incremental_remap.F: sum_tracers :7998, check_tracer_conservation :8126,
  tracer_local_min_max :8268, check_tracer_monotonicity :8416) -- ir_set_checks / ir_fetch_check_report /
  ir_fetch_conservation_sums.

Same two legs as tests/test_ir_parity.py: ``emulation`` (the shipped source compiled for the host, here) and ``cuda`` (the
shipped library on a B200, in a child process with a time limit).  The transported fields must be bit-identical with
and without the checks; the sums are held to 1e-13 relative (the device adds in a fixed tree, the oracle serially like
the reference); the reports must name the same first violation.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ir
from mpas_seaice_b200 import ir_host
from test_oracle_ir import case, smooth_divergent_velocity, _random_state
from test_ir_parity import _emulation_library, clone, CUDA_LEG_ENABLED

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LEGS = [
    pytest.param("emulation", id="emulation"),
    pytest.param("cuda", id="cuda", marks=[
        pytest.mark.gpu,
        pytest.mark.skipif(not CUDA_LEG_ENABLED, reason="runs in the child process of test_cuda_leg_in_a_child_process")]),
]


@pytest.fixture(params=LEGS)
def lib_path(request):
    if request.param == "emulation":
        return _emulation_library()
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return ir_host.LIB_PATH


@pytest.mark.gpu
def test_cuda_leg_in_a_child_process():
    """Every `cuda` case of this file in a process of its own with a time limit."""
    if CUDA_LEG_ENABLED:
        pytest.skip("this IS the child process")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, IR_B200_CUDA_LEG="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", "cuda and not child_process",
                        "-p", "no:cacheprovider", os.path.abspath(__file__)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


def _solver(kind, lib_path, n_categories, n_cells_solve=None):
    mesh, irf, geom = case(kind)
    return ir_host.IrTransport(mesh, irf, geom, n_categories, n_cells_solve=n_cells_solve, lib_path=lib_path)


def _assert_sums_close(d_ref, solver, tracers):
    for t, tr in enumerate(tracers):
        si, sf = solver.conservation_sums(t, tr.array.shape[2])
        for mine, ref in ((si, d_ref["sumInit"][t]), (sf, d_ref["sumFinal"][t])):
            assert np.all(np.abs(mine - ref) <= 1e-13 * np.abs(ref).max() + 1e-300), tr.name


@pytest.mark.parametrize("kind", ["hex16", "ico3"])
def test_checks_pass_and_leave_the_fields_alone(kind, lib_path):
    """Voronoi meshes (vertexDegree 3), divergent flow, full tracer hierarchy, three steps: both checks pass on the
    oracle (the reference's in-place extension and the order-independent one) and on the device, the sums agree, and
    the transported fields are bit-identical to a run without the checks and to the oracle."""
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    tracers = _random_state(mesh, np.random.default_rng(21))
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, ref_inplace, dev, plain = clone(tracers), clone(tracers), clone(tracers), clone(tracers)
    nK = tracers[0].array.shape[1]
    a, b = _solver(kind, lib_path, nK), _solver(kind, lib_path, nK)
    try:
        a.set_tracers(dev)
        a.set_checks(conservation=1, monotonicity=1)
        b.set_tracers(plain)
        for _ in range(3):
            d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, conservation_check=1, monotonicity_check=2)
            d_in = ir.run(mesh, irf, geom, ref_inplace, u, v, 3600.0, conservation_check=1, monotonicity_check=1)
            assert d_ref["error"] == 0 and d_in["error"] == 0
            assert a.run(dev, u, v, 3600.0) == 0
            assert b.run(plain, u, v, 3600.0) == 0
            rep = a.check_report()
            assert rep["conservationViolated"] == 0 and rep["monotonicityViolated"] == 0
            _assert_sums_close(d_ref, a, dev)
            # conservation itself: every mass * tracer product to round-off
            for i, f in zip(d_ref["sumInit"], d_ref["sumFinal"]):
                assert np.all(np.abs(f - i) <= 1e-13 * np.abs(i).max())
            for x, y, z in zip(ref, dev, plain):
                assert np.array_equal(x.array[:nC], y.array[:nC]), x.name
                assert np.array_equal(y.array[:nC], z.array[:nC]), x.name
    finally:
        a.destroy()
        b.destroy()


@pytest.mark.parametrize("kind", ["quad16", "band48"])
def test_monotonicity_report_matches_oracle(kind, lib_path):
    """Quadrilateral meshes (vertexDegree 4): ice reaches a cell from its diagonal neighbours, two edge rings away, and
    the limiter bounds their reconstructions by a third ring, so the reference's two-ring test can fire on a legitimate
    step.  Whether it does, and the first violation it reports (tracer, layer, category, cell, value, bound), must be
    the oracle's; the reference's in-place extension is never stricter than the order-independent one."""
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    tracers = _random_state(mesh, np.random.default_rng(21))
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, ref_inplace, dev = clone(tracers), clone(tracers), clone(tracers)
    s = _solver(kind, lib_path, tracers[0].array.shape[1])
    fired = 0
    try:
        s.set_tracers(dev)
        s.set_checks(conservation=1, monotonicity=1)
        for _ in range(3):
            d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, check=False, conservation_check=1, monotonicity_check=2)
            d_in = ir.run(mesh, irf, geom, ref_inplace, u, v, 3600.0, check=False, conservation_check=1, monotonicity_check=1)
            rc = s.run(dev, u, v, 3600.0, check=False)
            rep = s.check_report()
            assert d_ref["error"] in (0, 10) and rc == (ir_host.IR_ERR_MONOTONICITY if d_ref["error"] == 10 else 0)
            assert rep["conservationViolated"] == 0
            assert [rep["monotonicityViolated"], rep["monoTracer"], rep["monoLayer"], rep["monoCategory"], rep["monoCell"]] == \
                list(d_ref["monoErr"])
            if d_ref["error"] == 10:
                fired += 1
                assert (rep["newValue"], rep["bound"], rep["tolerance"]) == tuple(d_ref["monoVal"])
                assert b"monotonicity violation" in s._L.ir_last_error_string()
            else:
                assert d_in["error"] == 0            # in-place bounds contain the order-independent ones
            for x, y in zip(ref, dev):
                assert np.array_equal(x.array[:nC], y.array[:nC]), x.name
    finally:
        s.destroy()
    assert fired >= 1


def test_conservation_report_on_a_partial_block(lib_path):
    """A block that owns only part of its cells is not closed: ice leaves through the boundary of the owned region, and
    the local sums change.  conservation = 1 reports the first (tracer, category, layer) whose sum moved by more than
    1e-11 relative, the oracle's; conservation = 2 hands the same sums to the host (which would add the ranks) and
    returns IR_OK."""
    kind = "hex16"
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    nCS = (2 * nC) // 3
    tracers = _random_state(mesh, np.random.default_rng(5))
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, dev, dev2 = clone(tracers), clone(tracers), clone(tracers)
    nK = tracers[0].array.shape[1]
    a, b = _solver(kind, lib_path, nK, n_cells_solve=nCS), _solver(kind, lib_path, nK, n_cells_solve=nCS)
    try:
        d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, n_cells_solve=nCS, check=False, conservation_check=1,
                       monotonicity_check=2)
        assert d_ref["error"] == 9 and d_ref["consErr"][0] == 1
        a.set_tracers(dev)
        a.set_checks(conservation=1, monotonicity=1)
        assert a.run(dev, u, v, 3600.0, check=False) == ir_host.IR_ERR_CONSERVATION
        rep = a.check_report()
        assert [rep["conservationViolated"], rep["consTracer"], rep["consCategory"], rep["consLayer"]] == list(d_ref["consErr"])
        assert rep["monotonicityViolated"] == 0       # the reference aborts before the monotonicity check
        t, k, l = rep["consTracer"], rep["consCategory"] - 1, rep["consLayer"] - 1
        assert abs(rep["sumInit"] - d_ref["sumInit"][t][k, l]) <= 1e-13 * abs(d_ref["sumInit"][t][k, l])
        assert abs(rep["sumFinal"] - d_ref["sumFinal"][t][k, l]) <= 1e-13 * abs(d_ref["sumFinal"][t][k, l])
        _assert_sums_close(d_ref, a, dev)
        b.set_tracers(dev2)
        b.set_checks(conservation=2, monotonicity=0)
        assert b.run(dev2, u, v, 3600.0) == 0
        _assert_sums_close(d_ref, b, dev2)
        for x, y, z in zip(ref, dev, dev2):
            assert np.array_equal(x.array[:nC], y.array[:nC]) and np.array_equal(y.array[:nC], z.array[:nC]), x.name
    finally:
        a.destroy()
        b.destroy()


def test_checks_call_order_and_arguments(lib_path):
    kind = "hex12"
    mesh, irf, geom = case(kind)
    tracers = _random_state(mesh, np.random.default_rng(1), n_cat=2, n_ice=2, n_snow=1)
    u, v = smooth_divergent_velocity(mesh, geom)
    s = _solver(kind, lib_path, 2)
    try:
        with pytest.raises(ir_host.IrError):
            s.set_checks(conservation=3)
        with pytest.raises(ir_host.IrError):
            s.set_checks(monotonicity=2)
        s.set_tracers(tracers)
        with pytest.raises(ir_host.IrError):
            s.conservation_sums(0, 1)             # no run with the check on yet
        s.set_checks(conservation=1, monotonicity=1)
        assert s.run(tracers, u, v, 3600.0) == 0
        s.conservation_sums(0, 1)
        n_with = s.launch_count()
        s.set_checks(0, 0)                        # off again: the default five kernels (+ a layout kernel per tracer each way)
        assert s.run(tracers, u, v, 3600.0) == 0
        assert s.launch_count() - n_with == 5 + 2 * len(tracers)
        s.set_tracers(tracers)                    # a new tracer table forgets the sums
        with pytest.raises(ir_host.IrError):
            s.conservation_sums(0, 1)
    finally:
        s.destroy()
