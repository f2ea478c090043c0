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


# ----------------------------------------------------------------------------------------------------------------------
# seaice_normal_vectors (src/shared/mpas_seaice_mesh.F:703-846): oracle/upwind_oracle.c and ir_normal_vectors
# ----------------------------------------------------------------------------------------------------------------------
from oracle import upwind                                             # noqa: E402
from mpas_seaice_b200 import variational_init, weakmesh              # noqa: E402


@pytest.fixture(params=LEGS)
def leg(request):
    if request.param == "emulation":
        return "emulation", _emulation_library()
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return "cuda", ir_host.LIB_PATH


def _vertex_midpoint_edges(mesh, voe, eov):
    """x/y/zEdge as weakmesh.py takes them: the midpoint of the edge's vertices, pushed back onto the sphere."""
    nE = mesh.nEdges
    a, b = voe[:nE, 0] - 1, voe[:nE, 1] - 1
    p = [0.5 * (getattr(mesh, n)[a] + getattr(mesh, n)[b]) for n in ("xVertex", "yVertex", "zVertex")]
    if mesh.on_a_sphere:
        s = np.sqrt(p[0] ** 2 + p[1] ** 2 + p[2] ** 2) / mesh.sphere_radius
        p = [q / s for q in p]
    out = {}
    for name, q in zip(("xEdge", "yEdge", "zEdge"), p):
        out[name] = np.zeros(nE + 1)
        out[name][:nE] = q
    out.update(verticesOnEdge=voe, edgesOnVertex=eov)
    return out


def _assert_normals_close(got, ref, exact):
    """exact: bit for bit.  Otherwise first component to 1e-13 and the second, sign(n3) * sqrt(1 - n1**2), to what the
    first one's round-off allows: d(n2) = -n1 d(n1) / n2, i.e. |d(n2)| * max(|n2|, sqrt(eps)) <= 1e-13."""
    for key in ref:
        if exact:
            assert np.array_equal(got[key], ref[key]), key
        elif key.startswith("lat"):
            assert np.abs(got[key] - ref[key]).max() <= 1e-14, key
        else:
            d = np.abs(got[key] - ref[key])
            assert d[..., 0].max() <= 1e-13, key
            assert (d[..., 1] * np.maximum(np.abs(ref[key][..., 1]), 1.5e-8)).max() <= 1e-13, key


@pytest.mark.parametrize("kind", ["hex12", "quad10", "ico3", "band48"])
def test_oracle_normal_vectors_against_the_vectorised_restatement(kind):
    """oracle/upwind_oracle.c (loop for loop after mesh.F) against mpas-seaice_b200/weakmesh.py (numpy, written from
    the geometry: rotate to the equator of the cell, tangent x position, eastward component): two independent readings
    of seaice_normal_vectors with removeMetricTerms = .true. as the weak operators call it (weak.F:87-96)."""
    mesh, irf, _ = case(kind)
    wf = weakmesh.weak_fields(mesh)
    edges = _vertex_midpoint_edges(mesh, wf["verticesOnEdge"], wf["edgesOnVertex"])
    iv = variational_init.interior_vertex(mesh)
    o = upwind.normal_vectors(mesh, edges, iv, rotate=True, remove_metric_terms=True)
    _assert_normals_close(o, {k: wf[k] for k in o}, exact=False)
    n = o["normalVectorPolygon"][:mesh.nCells]
    used = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:mesh.nCells, None]
    assert np.abs((n ** 2).sum(axis=2) - 1.0)[used].max() < 1e-14          # unit vectors
    if not mesh.on_a_sphere:
        # a closed polygon: the side normals weighted by the side lengths add up to zero
        dv = mesh.dvEdge[mesh.edgesOnCell[:mesh.nCells] - 1]
        closed = (n * (dv * used)[:, :, None]).sum(axis=1)
        assert np.abs(closed).max() < 1e-9 * mesh.dvEdge[:mesh.nEdges].max()


@pytest.mark.parametrize("remove_metric_terms", [True, False])
@pytest.mark.parametrize("kind", ["hex12", "quad10", "ico3", "band48"])
def test_normal_vectors_match_oracle(kind, remove_metric_terms, leg):
    """ir_normal_vectors against the oracle, as the weak operators (removeMetricTerms) and as the upwind transport (not)
    call it.  Emulation: the same libm, bit for bit.  Device: CUDA's sin / cos / asin / atan2 are 1-2 ulp functions, so
    to round-off, with the second component held to what its formula allows (see _assert_normals_close)."""
    which, lib = leg
    mesh, irf, _ = case(kind)
    iv = variational_init.interior_vertex(mesh)
    ref = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=remove_metric_terms)
    got = ir_host.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=remove_metric_terms, lib_path=lib)
    _assert_normals_close(got, ref, exact=(which == "emulation"))
    ref_p = upwind.normal_vectors(mesh, irf, iv, rotate=False, remove_metric_terms=remove_metric_terms, triangles=False)
    got_p = ir_host.normal_vectors(mesh, irf, iv, rotate=False, remove_metric_terms=remove_metric_terms, triangles=False,
                                   lib_path=lib)
    assert set(got_p) == {"normalVectorPolygon", "latCellRotated"}
    _assert_normals_close(got_p, ref_p, exact=(which == "emulation"))


# ----------------------------------------------------------------------------------------------------------------------
# config_advection_type = 'upwind' (src/shared/mpas_seaice_advection_upwind.F): oracle/upwind_oracle.c and ir_run_upwind
# ----------------------------------------------------------------------------------------------------------------------

def _upwind_setup(kind):
    mesh, irf, geom = case(kind)
    iv = variational_init.interior_vertex(mesh)
    # seaice_init_advection_upwind (:122): polygons only, removeMetricTerms = .false.
    nve = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    return mesh, irf, geom, ir_host.interior_edge(mesh), nve


def _upwind_state(mesh, rng, n_cat=3, table="physical", ice_free=0.3):
    """area, ice volume, snow volume, surface temperature as (nCells+1, nCategories) arrays.  ``physical``: every tracer
    rides on the area.  ``reference``: define_tracer_connectivities as written (:160-165): area <- surfaceTemperature <-
    iceVolumeCategory <- snowVolumeCategory."""
    nC = mesh.nCells
    a = np.zeros((nC + 1, n_cat))
    a[:nC] = rng.uniform(0.05, 0.3, size=(nC, n_cat))
    a[:nC][rng.uniform(size=(nC, n_cat)) < ice_free] = 0.0
    vol, snow, tsfc = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    vol[:nC] = a[:nC] * rng.uniform(0.5, 3.0, size=(nC, n_cat))
    snow[:nC] = a[:nC] * rng.uniform(0.0, 0.3, size=(nC, n_cat))
    tsfc[:nC] = np.where(a[:nC] > 0, rng.uniform(-20.0, -1.0, size=(nC, n_cat)), 0.0)
    if table == "physical":
        return [upwind.Var("iceAreaCategory", a), upwind.Var("iceVolumeCategory", vol, 0, True),
                upwind.Var("snowVolumeCategory", snow, 0, True), upwind.Var("surfaceTemperature", tsfc, 0)]
    return [upwind.Var("iceAreaCategory", a), upwind.Var("surfaceTemperature", tsfc, 0),
            upwind.Var("iceVolumeCategory", vol, 1, True), upwind.Var("snowVolumeCategory", snow, 2, True)]


def _clone_vars(variables):
    return [upwind.Var(x.name, x.array.copy(), x.parent, x.volume_like, x.child_minimum) for x in variables]


def test_upwind_oracle_conserves_and_keeps_uniform_tracers():
    """Planar mesh, no flux through the boundary (its edges are not interior): the upwind step conserves the total area
    and the ice / snow volumes to round-off, a uniform thickness and a uniform temperature stay uniform wherever ice is
    left, and the new area of an interior cell is the closed-form donor-cell update."""
    mesh, irf, geom, interior, nve = _upwind_setup("quad16")
    nC = mesh.nCells
    rng = np.random.default_rng(3)
    var = _upwind_state(mesh, rng, n_cat=2)
    a0 = var[0].array.copy()
    var[1].array[:] = a0 * 1.7           # uniform thickness
    var[3].array[:] = np.where(a0 > 0, -5.0, 0.0)
    u0 = 0.04
    from test_oracle_ir import uniform_velocity
    u, v = uniform_velocity(mesh, u0, 0.0)
    dt = 3600.0
    before = [(x.array[:nC] * mesh.areaCell[:nC, None]).sum(axis=0) for x in var[:3]]
    d = upwind.run(mesh, irf["verticesOnEdge"], interior, nve, var, u, v, dt, diagnostics=True)
    after = [(x.array[:nC] * mesh.areaCell[:nC, None]).sum(axis=0) for x in var[:3]]
    for b, f in zip(before, after):
        assert np.all(np.abs(f - b) <= 1e-13 * np.abs(b))
    a1 = var[0].array
    ice = a1[:nC] > 1e-11
    assert np.abs(var[1].array[:nC][ice] / a1[:nC][ice] - 1.7).max() < 1e-13
    assert np.abs(var[3].array[:nC][ice] + 5.0).max() < 1e-13
    # donor cell: a_i <- a_i - (u dt / dx) (a_i - a_west) on cells whose four edges are interior and whose neighbourhood holds ice
    dx = 1000.0
    west = {}
    for c in range(nC):
        for j in range(mesh.nEdgesOnCell[c]):
            nb = mesh.cellsOnCell[c, j] - 1
            if nb < nC and abs(mesh.yCell[nb] - mesh.yCell[c]) < 1e-6 and mesh.xCell[nb] < mesh.xCell[c]:
                west[c] = nb
    from test_oracle_ir import inner_cells
    inner = np.nonzero(inner_cells(mesh, 1))[0]
    checked = 0
    for c in inner:
        w = west[c]
        if a0[c, 0] > 1e-11 and a0[w, 0] > 1e-11 and a0[[n - 1 for n in mesh.cellsOnCell[c, :4]], 0].min() > 1e-11:
            expect = a0[c, 0] - (u0 * dt / dx) * (a0[c, 0] - a0[w, 0])
            assert abs(a1[c, 0] - expect) < 1e-14
            checked += 1
    assert checked > 20
    assert np.abs(d["edgeVelocity"][:mesh.nEdges]).max() <= u0 * (1 + 1e-14)


def test_upwind_oracle_with_the_reference_table_zeroes_the_volumes():
    """define_tracer_connectivities as written chains iceVolumeCategory under surfaceTemperature (:162-164): the volume's
    "mass" is area * temperature, which is negative for ice below the melting point, `parentTracerNew > 0` (:1217) never
    holds and the volumes come back zero.  Recorded here as the as-executed behaviour of the reference's table; the
    physically meaningful table is the caller's choice (ir_upwind_var.parent)."""
    mesh, irf, geom, interior, nve = _upwind_setup("hex12")
    var = _upwind_state(mesh, np.random.default_rng(4), n_cat=2, table="reference")
    u, v = smooth_divergent_velocity(mesh, geom)
    a0 = var[0].array.copy()
    upwind.run(mesh, irf["verticesOnEdge"], interior, nve, var, u, v, 3600.0)
    assert np.all(var[2].array == 0.0) and np.all(var[3].array == 0.0)
    assert np.abs(var[0].array[:mesh.nCells] - a0[:mesh.nCells]).max() > 1e-6     # the area itself is transported


@pytest.mark.parametrize("table", ["physical", "reference"])
@pytest.mark.parametrize("kind", ["hex16", "quad16", "ico3", "band48"])
def test_upwind_matches_oracle(kind, table, lib_path):
    """ir_run_upwind against the oracle, bit for bit: every cell (halo and the extra slot included), the edge fluxes of
    every variable and the edge velocity, three steps, divergent flow, ice-free cells, a block that owns two thirds of
    its cells."""
    mesh, irf, geom, interior, nve = _upwind_setup(kind)
    nC = mesh.nCells
    nCS = (2 * nC) // 3
    var = _upwind_state(mesh, np.random.default_rng(11), table=table)
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, dev = _clone_vars(var), _clone_vars(var)
    s = _solver(kind, lib_path, var[0].array.shape[1], n_cells_solve=nCS)
    try:
        s.set_upwind_mesh(interior, mesh.dvEdge, nve)
        for _ in range(3):
            d = upwind.run(mesh, irf["verticesOnEdge"], interior, nve, ref, u, v, 3600.0, n_cells_solve=nCS, diagnostics=True)
            s.run_upwind(dev, u, v, 3600.0)
            for i, (x, y) in enumerate(zip(ref, dev)):
                assert np.array_equal(x.array, y.array), x.name
                flux, vel = s.upwind_fluxes(i)
                assert np.array_equal(flux[:mesh.nEdges], d["edgeFlux"][i][:mesh.nEdges]), x.name
                assert np.array_equal(vel[:mesh.nEdges], d["edgeVelocity"][:mesh.nEdges])
            # what a host does between steps: halo cells back to physical values (here: owned cells' rule applied by the oracle)
            for x, y in zip(ref, dev):
                y.array[:] = x.array
        assert np.abs(ref[0].array[:nCS] - var[0].array[:nCS]).max() > 1e-6
    finally:
        s.destroy()


def test_upwind_call_order_and_arguments(lib_path):
    mesh, irf, geom, interior, nve = _upwind_setup("hex12")
    var = _upwind_state(mesh, np.random.default_rng(2), n_cat=2)
    u, v = smooth_divergent_velocity(mesh, geom)
    s = _solver("hex12", lib_path, 2)
    try:
        with pytest.raises(ir_host.IrError) as e:
            s.run_upwind(var, u, v, 3600.0)                 # before ir_set_upwind_mesh
        assert e.value.code == ir_host.IR_ERR_STATE
        with pytest.raises(ir_host.IrError):
            s.upwind_fluxes(0)
        s.set_upwind_mesh(interior, mesh.dvEdge, nve)
        bad = _clone_vars(var)
        bad[1].parent = 2                                    # a parent after its child
        with pytest.raises(ir_host.IrError) as e:
            s.run_upwind(bad, u, v, 3600.0)
        assert e.value.code == ir_host.IR_ERR_ARGUMENT
        bad = _clone_vars(var)
        bad[0].parent = 0
        with pytest.raises(ir_host.IrError):
            s.run_upwind(bad, u, v, 3600.0)
        s.run_upwind(var, u, v, 3600.0)
        s.run_upwind(var[:2], u, v, 3600.0)                  # another table: the state is reallocated
        with pytest.raises(ir_host.IrError):
            s.upwind_fluxes(2)
    finally:
        s.destroy()


# ----------------------------------------------------------------------------------------------------------------------
# Reference-executed fixtures (tests/golden/options/*.npz): outputs of the reference's own Fortran source -- mesh.F's
# seaice_normal_vectors and advection_upwind.F's define_tracer_connectivities / seaice_run_advection_upwind -- interpreted
# by tests/golden/fortran_subset.py (generator: tests/golden/make_reference_executed_golden.py).
# ----------------------------------------------------------------------------------------------------------------------
import ast      # noqa: E402
import glob     # noqa: E402

from mpas_seaice_b200 import irmesh, meshgen     # noqa: E402

_OPT = os.path.join(ROOT, "tests", "golden", "options")
NORMAL_FILES = sorted(glob.glob(os.path.join(_OPT, "refexec_normals_*.npz")))
UPWIND_FILES = sorted(glob.glob(os.path.join(_OPT, "refexec_upwind_*.npz")))


def _fixture_mesh(z):
    spec = ast.literal_eval(str(z["spec"]))
    mesh = getattr(meshgen, spec[0])(*spec[1:])
    for k in z.files:
        if k.startswith("mesh_"):
            assert np.array_equal(mesh[k[5:]], z[k]), "meshgen no longer produces the fixture's mesh"
    return mesh, irmesh.ir_fields(mesh)


def test_reference_executed_option_fixtures_exist():
    assert len(NORMAL_FILES) == 8 and len(UPWIND_FILES) == 3
    seen = set()
    for f in NORMAL_FILES + UPWIND_FILES:
        prov = str(np.load(f)["provenance"])
        assert "interpreting the reference's Fortran source" in prov
        seen |= {w.strip() for w in prov.split(":", 1)[1].split(",")}
    for name in ("seaice_normal_vectors", "normal_vectors_planar_polygon", "normal_vectors_planar_triangle",
                 "normal_vectors_spherical_polygon_metric", "normal_vectors_spherical_triangle_metric",
                 "define_tracer_connectivities", "seaice_run_advection_upwind", "edge_from_vertex_velocity",
                 "prepare_none_parent_tracer", "upwind_tendencies", "run_advection_subvariable", "scale_tracers_back_3d"):
        assert name in seen, name
    # the table the reference's own define_tracer_connectivities builds (:160-165), read back from the interpreter
    table = ast.literal_eval(str(np.load(UPWIND_FILES[0])["table"]))
    assert table == [("iceAreaCategory", "none"), ("surfaceTemperature", "iceAreaCategory"),
                     ("iceVolumeCategory", "surfaceTemperature"), ("snowVolumeCategory", "iceVolumeCategory")]


@pytest.mark.parametrize("path", NORMAL_FILES, ids=[os.path.basename(f)[16:-4] for f in NORMAL_FILES])
def test_normal_vectors_reproduce_the_reference_executed_arrays(path, leg):
    """oracle/upwind_oracle.c: bit for bit.  ir_normal_vectors: bit for bit under emulation (same libm), to round-off on
    the device (CUDA's trigonometric functions are 1-2 ulp; tolerance model in _assert_normals_close)."""
    which, lib = leg
    z = np.load(path)
    mesh, irf = _fixture_mesh(z)
    iv = variational_init.interior_vertex(mesh)
    rm = bool(z["remove_metric_terms"])
    want = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    _assert_normals_close(upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=rm), want, exact=True)
    got = ir_host.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=rm, lib_path=lib)
    _assert_normals_close(got, want, exact=(which == "emulation"))


@pytest.mark.parametrize("path", UPWIND_FILES, ids=[os.path.basename(f)[15:-4] for f in UPWIND_FILES])
def test_upwind_reproduces_the_reference_executed_steps(path, lib_path):
    """Two steps of seaice_run_advection_upwind as the reference's source executes them, its own connectivity table:
    the oracle and ir_run_upwind reproduce the four advected arrays bit for bit (every cell; the volumes come out zero,
    see test_upwind_oracle_with_the_reference_table_zeroes_the_volumes).  The fixture also records what the reference
    does to the tracers that are NOT in its table: the time-level shift leaves them zero."""
    z = np.load(path)
    mesh, irf = _fixture_mesh(z)
    names = ("iceAreaCategory", "surfaceTemperature", "iceVolumeCategory", "snowVolumeCategory")
    parents = {"iceAreaCategory": None, "surfaceTemperature": 0, "iceVolumeCategory": 1, "snowVolumeCategory": 2}
    make = lambda: [upwind.Var(n, z["in_" + n].copy(), parents[n], n.endswith("VolumeCategory")) for n in names]
    ref, dev = make(), make()
    u, v, nve, interior = z["in_uVelocity"], z["in_vVelocity"], z["in_normalVectorEdge"], z["in_interiorEdge"]
    s = ir_host.IrTransport(mesh, irf, case_geometry(mesh, irf), ref[0].array.shape[1], lib_path=lib_path)
    try:
        s.set_upwind_mesh(interior, mesh.dvEdge, nve)
        for step in range(1, int(z["nsteps"]) + 1):
            upwind.run(mesh, irf["verticesOnEdge"], interior, nve, ref, u, v, float(z["dt"]))
            s.run_upwind(dev, u, v, float(z["dt"]))
            for x, y in zip(ref, dev):
                want = z["out%d_%s" % (step, x.name)]
                assert np.array_equal(x.array, want), ("oracle", x.name, step)
                assert np.array_equal(y.array, want), ("device", x.name, step)
            for n in ("iceEnthalpy", "iceSalinity", "snowEnthalpy"):
                assert np.all(z["out%d_%s" % (step, n)] == 0.0)
        assert np.abs(z["out2_iceAreaCategory"] - z["in_iceAreaCategory"]).max() > 1e-6
    finally:
        s.destroy()


def case_geometry(mesh, irf):
    return ir.init_geometry(mesh, irf)


def test_option_kernels_do_not_depend_on_the_thread_order():
    """The emulation runs blocks and threads first-to-last or last-to-first (IR_EMU_ORDER): the checks (sums, local
    extremes, the first violation found through atomicMin), the upwind step (edge kernels feeding cell kernels, the in-place
    rescaling of the old time level) and the normal vectors give the same bits either way -- no thread of a launch depends
    on another one of the same launch."""
    lib = _emulation_library()
    mesh, irf, geom, interior, nve = _upwind_setup("quad16")
    iv = variational_init.interior_vertex(mesh)
    tracers = _random_state(mesh, np.random.default_rng(21))
    var = _upwind_state(mesh, np.random.default_rng(11))
    u, v = smooth_divergent_velocity(mesh, geom)
    results = []
    for order in ("forward", "reverse"):
        os.environ["IR_EMU_ORDER"] = order
        try:
            dev, uw = clone(tracers), _clone_vars(var)
            s = _solver("quad16", lib, tracers[0].array.shape[1], n_cells_solve=(2 * mesh.nCells) // 3)
            try:
                s.set_tracers(dev)
                s.set_checks(2, 1)
                rc = s.run(dev, u, v, 3600.0, check=False)
                rep = s.check_report()
                sums = [s.conservation_sums(i, t.array.shape[2]) for i, t in enumerate(dev)]
                s.set_upwind_mesh(interior, mesh.dvEdge, nve)
                for _ in range(2):
                    s.run_upwind(uw, u, v, 3600.0)
                flux = s.upwind_fluxes(1)
            finally:
                s.destroy()
            nv = ir_host.normal_vectors(mesh, irf, iv, lib_path=lib)
        finally:
            os.environ.pop("IR_EMU_ORDER", None)
        results.append((rc, rep, sums, dev, uw, flux, nv))
    a, b = results
    assert a[0] == b[0] == ir_host.IR_ERR_MONOTONICITY and a[1] == b[1]
    for (si0, sf0), (si1, sf1) in zip(a[2], b[2]):
        assert np.array_equal(si0, si1) and np.array_equal(sf0, sf1)
    for x, y in zip(a[3], b[3]):
        assert np.array_equal(x.array, y.array), x.name
    for x, y in zip(a[4], b[4]):
        assert np.array_equal(x.array, y.array), x.name
    assert np.array_equal(a[5][0], b[5][0]) and np.array_equal(a[5][1], b[5][1])
    for k in a[6]:
        assert np.array_equal(a[6][k], b[6][k]), k
