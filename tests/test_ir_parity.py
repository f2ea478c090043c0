"""Parity of the incremental-remapping kernels (mpas-seaice_b200/csrc/ir_kernels.cu, C ABI include/ir_b200.h) with
the oracle (oracle/ir_oracle.c): bit-exact -- FP64 work done in the reference's operation order with contraction
off on both sides.

Two legs run the same cases through the same C ABI and ctypes host:

* ``cuda`` (@gpu): the shipped library on a B200, in a child process (see test_cuda_leg_in_a_child_process).
  First passed on a B200 in round 1's driver GPUTEST; a difference fails the suite.
* ``emulation`` (CPU): the SAME source file compiled with g++ against tests/emu/cuda_runtime.h, which runs every
  kernel thread sequentially.  It checks the kernels' logic here, where there is no GPU; it is test infrastructure,
  built under tests/_build/, and never shipped or loaded by the product.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import ir
from mpas_seaice_b200 import ir_host
from test_oracle_ir import case, smooth_divergent_velocity, uniform_velocity, _random_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# IR_B200_EMU_SRC: run the emulation leg on another source file with the same ABI (a kernel variant under development)
SRC = os.environ.get("IR_B200_EMU_SRC") or os.path.join(ROOT, "mpas-seaice_b200", "csrc", "ir_kernels.cu")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(ROOT, "tests", "_build", "libir_emu%s.so" % ("" if "IR_B200_EMU_SRC" not in os.environ else
                                                                   "_" + os.path.splitext(os.path.basename(SRC))[0]))


def _emulation_library():
    deps = [SRC, os.path.join(os.path.dirname(SRC), "ir_upwind.cuh"), os.path.join(EMU_DIR, "cuda_runtime.h"),
            os.path.join(ROOT, "include", "ir_b200.h")]
    deps = [p for p in deps if os.path.exists(p)]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(p) > os.path.getmtime(EMU_LIB) for p in deps):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        "-Wall", "-Wno-unknown-pragmas", "-I", EMU_DIR, "-o", EMU_LIB, SRC], check=True)
    return EMU_LIB


# The cuda leg runs in a child process (test_cuda_leg_in_a_child_process below) with a time limit: a fault there must
# not take the CUDA context -- or the process -- of the EVP tests with it.
CUDA_LEG_ENABLED = os.environ.get("IR_B200_CUDA_LEG") == "1"
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
    """Every `cuda` case of this file and of test_ir_blocks.py, in a process of its own with a time limit."""
    import sys
    if CUDA_LEG_ENABLED:
        pytest.skip("this IS the child process")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, IR_B200_CUDA_LEG="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", "cuda and not child_process",
                        "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_ir_parity.py"), os.path.join(ROOT, "tests", "test_ir_blocks.py")],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


def clone(tracers):
    return [ir.Tracer(t.name, t.array.copy(), t.parent, t.volume_like) for t in tracers]


def run_both(kind, tracers, u, v, dt, lib_path, steps=1, n_quad_points=6, rotate=False, check=True):
    mesh, irf, geom = case(kind)
    if rotate:
        geom = ir.init_geometry(mesh, irf, rotate=True)
    ref, dev = clone(tracers), clone(tracers)
    nK = tracers[0].array.shape[1]
    solver = ir_host.IrTransport(mesh, irf, geom, nK, n_quad_points=n_quad_points, rotate=rotate, lib_path=lib_path)
    try:
        solver.set_tracers(dev)
        codes = []
        for _ in range(steps):
            d_ref = ir.run(mesh, irf, geom, ref, u, v, dt, n_quad_points=n_quad_points, rotate=rotate, diagnostics=True,
                           check=False)
            rc = solver.run(dev, u, v, dt, check=False)
            codes.append((d_ref["error"], rc))
        d_dev = solver.diagnostics(tracers[0].array.shape[2])
        launches = solver.launch_count()
    finally:
        solver.destroy()
    return mesh, ref, dev, d_ref, d_dev, codes, launches


def assert_identical(mesh, ref, dev, d_ref, d_dev):
    nC = mesh.nCells
    for a, b in zip(ref, dev):
        assert np.array_equal(a.array[:nC], b.array[:nC]), a.name
    for key in ("maskEdge", "iCellTriangle", "triangleArea", "xTriangle", "yTriangle", "edgeFluxMass"):
        assert np.array_equal(d_ref[key], d_dev[key]), key


@pytest.mark.parametrize("kind", ["hex16", "quad16", "ico3", "band48"])
def test_full_hierarchy_matches_oracle(kind, lib_path):
    """area -> {volume -> {enthalpy, salinity (layers)}, snow volume -> snow enthalpy, surface temperature}, three
    categories, divergent flow, ice-free cells: three steps, every tracer and every departure triangle identical."""
    mesh, irf, geom = case(kind)
    rng = np.random.default_rng(21)
    tracers = _random_state(mesh, rng)
    u, v = smooth_divergent_velocity(mesh, geom)
    mesh, ref, dev, d_ref, d_dev, codes, launches = run_both(kind, tracers, u, v, 3600.0, lib_path, steps=3)
    assert all(c == (0, 0) for c in codes), codes
    assert_identical(mesh, ref, dev, d_ref, d_dev)
    assert any(np.any(a.array != t.array) for a, t in zip(ref, tracers))      # something moved
    assert launches > 0


@pytest.mark.parametrize("kind,vel", [("hex16", (0.05, 0.02)), ("hex16", (-0.04, -0.04)), ("quad16", (0.07, 0.0)),
                                      ("quad16", (-0.03, 0.06)), ("quad16", (0.0, -0.05))])
def test_uniform_flow_matches_oracle(kind, vel, lib_path):
    """Every configuration of find_departure_triangles: side triangles left / right, a quadrilateral on one side of
    the edge, the E5/E6 split on quads."""
    mesh, irf, geom = case(kind)
    nC = mesh.nCells
    tracers = ir.default_tracers(nC, 2)
    x, y = mesh.xCell[:nC], mesh.yCell[:nC]
    tracers[0].array[:nC, 0, 0] = 0.3 + 1e-5 * x + 2e-5 * y
    tracers[0].array[:nC, 1, 0] = 0.2 - 0.5e-5 * x + 1e-5 * y
    tracers[1].array[:nC] = tracers[0].array[:nC] * 2.0
    tracers[3].array[:nC] = -5.0
    u, v = uniform_velocity(mesh, *vel)
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both(kind, tracers, u, v, 3600.0, lib_path)
    assert codes == [(0, 0)]
    assert_identical(mesh, ref, dev, d_ref, d_dev)


def test_three_point_quadrature_and_rotated_grid_match_oracle(lib_path):
    """nQuadPoints = 3 (with the reference's x-for-y mid-point, :6598) on a plane, where it stays benign; the rotated
    Cartesian grid (config_rotate_cartesian_grid: gradient components swapped, :4365-4372) on the sphere."""
    mesh, irf, geom = case("hex16")
    rng = np.random.default_rng(4)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0, ice_free=0.0)
    u, v = uniform_velocity(mesh, 0.04, -0.03)
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both("hex16", tracers, u, v, 3600.0, lib_path, n_quad_points=3, check=False)
    assert codes == [(0, 0)]
    assert_identical(mesh, ref, dev, d_ref, d_dev)
    mesh, irf, geom = case("ico3")
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom, cfl=0.2)
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both("ico3", tracers, u, v, 3600.0, lib_path, rotate=True, check=False)
    assert codes == [(0, 0)]
    assert_identical(mesh, ref, dev, d_ref, d_dev)


def test_abort_conditions_are_reported_like_the_oracle(lib_path):
    """The 3-point quadrature on the sphere puts its points far outside the departure triangles (the mid-point slip,
    :6598): masses go negative, which the reference treats as fatal (:6895, :7465).  Oracle and device must both
    report it; departure triangles and mass fluxes, computed before the abort, are still identical."""
    mesh, irf, geom = case("ico3")
    rng = np.random.default_rng(4)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom, cfl=0.2)
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both("ico3", tracers, u, v, 3600.0, lib_path, n_quad_points=3, check=False)
    oracle_code, device_code = codes[0]
    assert oracle_code in (4, 5)
    assert device_code in (ir_host.IR_ERR_NEGATIVE_MASS_QP, ir_host.IR_ERR_NEGATIVE_MASS)
    for key in ("maskEdge", "iCellTriangle", "triangleArea", "xTriangle", "yTriangle", "edgeFluxMass"):
        assert np.array_equal(d_ref[key], d_dev[key]), key
    nC = mesh.nCells
    assert np.array_equal(ref[0].array[:nC], dev[0].array[:nC])      # the mass field was updated before the check


def test_deep_and_layered_hierarchies_match_oracle(lib_path):
    """Three parents (area -> volume -> brine fraction -> a layered child: the depth of mobileFraction /
    verticalAlgaeIce, incremental_remap_tracers.F:300-330) and a layered tracer under a layered parent
    (compute_barycenter_coordinates with 3-D parents, :4040-4170); mass * tracer products conserved, device =
    oracle."""
    mesh, irf, geom = case("ico3")
    nC = mesh.nCells
    rng = np.random.default_rng(33)

    def rnd(nl, lo, hi):
        a = np.zeros((nC + 1, 2, nl))
        a[:nC] = rng.uniform(lo, hi, (nC, 2, nl))
        return a
    area = rnd(1, 0.05, 0.45)
    area[:nC, :, 0][rng.uniform(size=(nC, 2)) < 0.2] = 0.0
    tracers = [ir.Tracer("iceAreaCategory", area),
               ir.Tracer("iceVolumeCategory", area * rnd(1, 0.5, 3.0), 0, True),
               ir.Tracer("layeredParent", rnd(2, 0.2, 1.0), 0),
               ir.Tracer("brineFraction", rnd(1, 0.3, 1.0), 1),
               ir.Tracer("layeredChild", rnd(2, 1.0, 5.0), 2),
               ir.Tracer("mobileFraction", rnd(2, 0.1, 0.9), 3)]
    A = mesh.areaCell[:nC, None, None]

    def products(tr):
        a, vol = tr[0].array[:nC], tr[1].array[:nC]
        return [(a * A).sum(), (vol * A).sum(), (a * tr[2].array[:nC] * A).sum(0), (vol * tr[3].array[:nC] * A).sum(),
                (a * tr[2].array[:nC] * tr[4].array[:nC] * A).sum(0), (vol * tr[3].array[:nC] * tr[5].array[:nC] * A).sum(0)]
    before = products(tracers)
    u, v = smooth_divergent_velocity(mesh, geom)
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both("ico3", tracers, u, v, 3600.0, lib_path, steps=2)
    assert all(c == (0, 0) for c in codes), codes
    assert_identical(mesh, ref, dev, d_ref, d_dev)
    for b, a in zip(before, products(ref)):
        assert np.allclose(a, b, rtol=5e-13, atol=0)


def test_emulated_kernels_do_not_depend_on_the_thread_order():
    """The emulation runs the threads of a launch sequentially; run last-to-first (IR_EMU_ORDER=reverse) the results
    must not change -- they would if a thread read what another thread of the same launch writes, which on the device
    is a race.  Full hierarchy, geometry init included."""
    lib = _emulation_library()
    mesh, irf, _ = case("ico3")
    rng = np.random.default_rng(21)
    tracers = _random_state(mesh, rng)
    results = []
    for order in ("forward", "reverse"):
        os.environ["IR_EMU_ORDER"] = order
        try:
            geom = ir_host.init_geometry(mesh, irf, lib_path=lib)
            u, v = smooth_divergent_velocity(mesh, geom)
            dev = clone(tracers)
            solver = ir_host.IrTransport(mesh, irf, geom, tracers[0].array.shape[1], lib_path=lib)
            try:
                solver.set_tracers(dev)
                for _ in range(2):
                    solver.run(dev, u, v, 3600.0)
                diag = solver.diagnostics()
            finally:
                solver.destroy()
        finally:
            os.environ.pop("IR_EMU_ORDER", None)
        results.append((geom, dev, diag))
    (g0, t0, d0), (g1, t1, d1) = results
    for k in ("xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap"):
        assert np.array_equal(g0[k], g1[k]), k
    for a, b in zip(t0, t1):
        assert np.array_equal(a.array, b.array), a.name
    for k in d0:
        assert np.array_equal(d0[k], d1[k]), k


def test_emulated_kernels_are_clean_under_address_sanitizer(tmp_path):
    """Device-memory bounds, checked where there is no device: the kernels compiled for the host with
    -fsanitize=address,undefined (no recovery), every "device" buffer a calloc of exactly the requested size, poisoned."""
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not installed")
    lib = str(tmp_path / "libir_emu_asan.so")
    subprocess.run(["g++", "-O1", "-g", "-fsanitize=address", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-ffp-contract=off", "-std=c++17",
                    "-fPIC", "-shared", "-x", "c++", "-Wno-unknown-pragmas", "-I", EMU_DIR, "-o", lib, SRC], check=True)
    # IR_EMU_POISON: fresh "device" memory holds 0xFF bytes, not zeros (cudaMalloc does not clear memory either)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", IR_EMU_POISON="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ir_asan_worker.py"), lib], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.stdout.count("ok") == 9


def test_work_fields_match_oracle(lib_path):
    """ir_fetch_tracer_field: the limited gradients of a child tracer and the mass fluxes, against the oracle's."""
    mesh, irf, geom = case("ico3")
    rng = np.random.default_rng(5)
    tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=0)
    u, v = smooth_divergent_velocity(mesh, geom)
    ref, dev = clone(tracers), clone(tracers)
    d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, diagnostics=True, grad_tracer=4)
    solver = ir_host.IrTransport(mesh, irf, geom, 2, lib_path=lib_path)
    try:
        solver.set_tracers(dev)
        solver.run(dev, u, v, 3600.0)
        nC = mesh.nCells
        assert np.array_equal(solver.tracer_field("xGrad", 4, 2)[:nC], d_ref["xGrad"][:nC])
        assert np.array_equal(solver.tracer_field("yGrad", 4, 2)[:nC], d_ref["yGrad"][:nC])
        assert np.array_equal(solver.tracer_field("edgeFlux", 0, 1)[:mesh.nEdges], d_ref["edgeFluxMass"])
        cen = solver.tracer_field("center", 0, 1)
        xg, yg = solver.tracer_field("xGrad", 0, 1), solver.tracer_field("yGrad", 0, 1)
        a = tracers[0].array
        want = a[:nC, :, 0] - xg[:nC, :, 0] * geom["geomAvg"]["x"][:nC, None] - yg[:nC, :, 0] * geom["geomAvg"]["y"][:nC, None]
        assert np.array_equal(cen[:nC, :, 0], want)
        # the centre of mass of the reconstructed area field (compute_barycenter_coordinates, one-field branch :4700)
        g = geom["geomAvg"]
        ice = (tracers[0].array[:nC, :, 0].sum(axis=1) > 0)[:, None] & (a[:nC, :, 0] > 0)
        xb = solver.tracer_field("xBarycenter", 0, 1)[:nC, :, 0]
        want = (cen[:nC, :, 0] * g["x"][:nC, None] + xg[:nC, :, 0] * g["xx"][:nC, None] + yg[:nC, :, 0] * g["xy"][:nC, None]) \
            / np.where(ice, a[:nC, :, 0], 1.0)
        assert np.allclose(xb[ice], want[ice], rtol=1e-14, atol=0)
        with pytest.raises(ir_host.IrError):
            solver.tracer_field("center", 9, 1)
        with pytest.raises(ir_host.IrError, match="children"):
            solver.tracer_field("xBarycenter", 3, 1)         # surfaceTemperature has no children
    finally:
        solver.destroy()


@pytest.mark.parametrize("kind", ["hex16", "quad16", "ico3", "band48"])
def test_random_states_and_velocities_match_oracle(kind, lib_path):
    """Fuzz: random ice cover, random CFL up to 0.6, velocity fields alternately smooth and random per vertex.  The
    rough fields reach the rare configurations -- two side triangles at one vertex (more than four departure triangles
    on a hexagonal edge: first found by this test), departure regions leaving their source cell (the reference's abort
    conditions).  Device and oracle must agree on the triangles and fluxes always, on every tracer when neither aborts,
    and on whether the step is fatal."""
    mesh, irf, geom = case(kind)
    nC, nV = mesh.nCells, mesh.nVertices
    solver = ir_host.IrTransport(mesh, irf, geom, 2, lib_path=lib_path)
    seen_many = False
    try:
        for seed in range(8):
            rng = np.random.default_rng(1000 + seed)
            tracers = _random_state(mesh, rng, n_cat=2, n_ice=2, n_snow=1, ice_free=rng.uniform(0, 0.6))
            cfl = rng.uniform(0.1, 0.6)
            if seed % 2 == 0:
                speed = cfl * geom["minLengthEdgesOnVertex"][:nV].min() / 3600.0
                u, v = np.zeros(nV + 1), np.zeros(nV + 1)
                u[:nV], v[:nV] = rng.uniform(-speed, speed, nV), rng.uniform(-speed, speed, nV)
            else:
                uu, vv = smooth_divergent_velocity(mesh, geom, cfl=cfl)
                ang = rng.uniform(0, 2 * np.pi)
                u, v = uu * np.cos(ang) - vv * np.sin(ang), uu * np.sin(ang) + vv * np.cos(ang)
            ref, dev = clone(tracers), clone(tracers)
            d_ref = ir.run(mesh, irf, geom, ref, u, v, 3600.0, check=False, diagnostics=True)
            solver.set_tracers(dev)
            rc = solver.run(dev, u, v, 3600.0, check=False)
            d_dev = solver.diagnostics()
            for key in d_dev:
                assert np.array_equal(d_ref[key], d_dev[key]), (seed, key)
            seen_many = seen_many or np.count_nonzero(d_ref["triangleArea"], axis=1).max() > 4
            assert (d_ref["error"] != 0) == (rc != 0), (seed, d_ref["error"], rc)
            if d_ref["error"] == 0:
                for a, b in zip(ref, dev):
                    assert np.array_equal(a.array[:nC], b.array[:nC]), (seed, a.name)
            elif d_ref["error"] == 4:        # flagged after the whole step ran: the fields are still comparable
                assert np.array_equal(ref[0].array[:nC], dev[0].array[:nC]), seed
    finally:
        solver.destroy()
    if kind == "ico3":
        assert seen_many                     # the case that needs the overflow path of k_fluxes is in the sample


def test_rotation_test_case_matches_oracle(lib_path):
    """The reference's advection test case (cosine bell, u = U cos(lat); create_ics.py:36-107) on the 2562-cell
    sphere with its twelve pentagons: ten steps, identical throughout."""
    import math
    from test_oracle_ir import _rotation_case
    mesh, irf, geom = case("ico4")
    nC, nV = mesh.nCells, mesh.nVertices
    p, field = _rotation_case(mesh, "cosine_bell")
    tracers = ir.default_tracers(nC, 1)
    tracers[0].array[:nC, 0, 0] = field(p)
    tracers[1].array[:nC, 0, 0] = tracers[0].array[:nC, 0, 0] * (1.0 + 0.5 * p[:, 2])
    u, v = np.zeros(nV + 1), np.zeros(nV + 1)
    u[:nV] = 2.0 * math.pi * 6371229.0 / (120.0 * 86400.0) * np.cos(mesh.latVertex[:nV])
    mesh, ref, dev, d_ref, d_dev, codes, _ = run_both("ico4", tracers, u, v, 6.0 * 3600.0, lib_path, steps=10)
    assert all(c == (0, 0) for c in codes)
    assert_identical(mesh, ref, dev, d_ref, d_dev)


@pytest.mark.parametrize("kind,rotate", [("hex12", False), ("quad10", False), ("ico3", False), ("ico4", False), ("ico3", True),
                                         ("band48", False), ("band48", True)])
def test_init_geometry_matches_oracle(kind, rotate, lib_path):
    """ir_init_geometry (the incremental_remap pool arrays for hosts without the Fortran init) against
    orc_ir_init_geometry: every array identical, on planar hexes and quads (with their boundary stencils), on the
    sphere, and with the rotated Cartesian grid."""
    mesh, irf, _ = case(kind)
    ref = ir.init_geometry(mesh, irf, rotate=rotate)
    got = ir_host.init_geometry(mesh, irf, rotate=rotate, lib_path=lib_path)
    nC, nE, nV = mesh.nCells, mesh.nEdges, mesh.nVertices
    for name in ("xVertexOnCell", "yVertexOnCell"):
        assert np.array_equal(ref[name][:nC], got[name][:nC]), name
    for name in ("remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap", "xVertexOnEdge", "yVertexOnEdge"):
        assert np.array_equal(ref[name][:nE], got[name][:nE]), name
    assert np.array_equal(ref["minLengthEdgesOnVertex"][:nV], got["minLengthEdgesOnVertex"][:nV])
    if mesh.on_a_sphere:
        assert np.array_equal(ref["transGlobalToCell"], got["transGlobalToCell"])
    for name in ir_host.GEOM_NAMES:
        assert np.array_equal(ref["geomAvg"][name][:nC], got["geomAvg"][name][:nC]), name


def test_init_geometry_with_halo_cells_and_bad_orientation(lib_path):
    """remapEdge covers the edges of OWNED cells only (nCellsSolve < nCells, :1235-1262); a mesh whose verticesOnEdge
    runs the wrong way round is refused like the reference refuses it (:1296)."""
    mesh, irf, _ = case("hex12")
    n_solve = mesh.nCells // 2
    ref = ir.init_geometry(mesh, irf, n_cells_solve=n_solve)
    got = ir_host.init_geometry(mesh, irf, n_cells_solve=n_solve, lib_path=lib_path)
    assert 0 < got["remapEdge"].sum() < case("hex12")[2]["remapEdge"].sum()
    for name in ("remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap", "xVertexOnEdge", "yVertexOnEdge"):
        assert np.array_equal(ref[name][:mesh.nEdges], got[name][:mesh.nEdges]), name
    bad = dict(irf)
    bad["verticesOnEdge"] = np.ascontiguousarray(irf["verticesOnEdge"][:, ::-1])
    with pytest.raises(ir_host.IrError) as e:
        ir_host.init_geometry(mesh, bad, lib_path=lib_path)
    assert e.value.code == ir_host.IR_ERR_MESH
    with pytest.raises(RuntimeError, match="edge orientation"):
        ir.init_geometry(mesh, bad)


def test_call_order_and_argument_errors(lib_path):
    mesh, irf, geom = case("hex12")
    solver = ir_host.IrTransport(mesh, irf, geom, 1, lib_path=lib_path)
    try:
        tracers = ir.default_tracers(mesh.nCells, 1)
        u, v = uniform_velocity(mesh, 0.0, 0.0)
        with pytest.raises(ir_host.IrError) as e:
            solver.run(tracers, u, v, 1.0)
        assert e.value.code == ir_host.IR_ERR_STATE
        bad = clone(tracers)
        bad[2].parent = 3                      # a child before its parent
        with pytest.raises(ir_host.IrError) as e:
            solver.set_tracers(bad)
        assert e.value.code == ir_host.IR_ERR_ARGUMENT
        solver.set_tracers(tracers)
        with pytest.raises(ir_host.IrError) as e:
            solver.run(tracers[:3], u, v, 1.0)
        assert e.value.code == ir_host.IR_ERR_ARGUMENT
        assert solver.run(tracers, u, v, 1.0) == ir_host.IR_OK
    finally:
        solver.destroy()
    with pytest.raises(ir_host.IrError):
        ir_host.IrTransport(mesh, irf, geom, 1, n_quad_points=4, lib_path=lib_path)


def test_shipped_library_exports_the_abi():
    """include/ir_b200.h against the built CUDA library (no compute: there is no device here)."""
    import ctypes as C
    import re
    header = open(os.path.join(ROOT, "include", "ir_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*(ir_[a-z_]+)\(", header, flags=re.M))
    assert declared == set(ir_host.EXPORTS)
    L = C.CDLL(ir_host.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Wextra", "-x", "c", os.path.join(ROOT, "include", "ir_b200.h")],
                   check=True)


# ----------------------------------------------------------------------------------------------------------------------
# Reference-executed fixtures (tests/golden/ir/*.npz): seaice_run_advection_incremental_remap with everything below it,
# interpreted from the reference's own Fortran source by tests/golden/fortran_subset.py (generator:
# tests/golden/make_reference_executed_golden.py).  The oracle and the kernels must reproduce every tracer bit for bit.
# ----------------------------------------------------------------------------------------------------------------------
import ast      # noqa: E402
import glob     # noqa: E402

REFEXEC_IR = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ir", "refexec_ir_*.npz")))


def test_reference_executed_transport_fixtures_exist():
    assert len(REFEXEC_IR) >= 6
    seen = set()
    for f in REFEXEC_IR:
        prov = str(np.load(f)["provenance"])
        assert "interpreting the reference's Fortran source" in prov
        seen |= {w.strip() for w in prov.split(":", 1)[1].split(",")}
    for name in ("seaice_run_advection_incremental_remap", "incremental_remap_block", "make_masks", "construct_linear_tracer_fields",
                 "compute_gradient_2d", "compute_gradient_3d", "limit_tracer_gradient_2d", "limit_tracer_gradient_3d",
                 "compute_barycenter_coordinates", "find_departure_points", "find_departure_triangles",
                 "shift_vertices_of_departure_triangle", "get_triangle_quadrature_points", "integrate_fluxes_over_triangles",
                 "compute_mass_tracer_products", "update_mass_and_tracers", "zap_small_mass", "volume_to_thickness",
                 "thickness_to_volume", "sum_tracers", "tracer_local_min_max", "check_tracer_conservation",
                 "check_tracer_monotonicity"):
        assert name in seen, name


@pytest.mark.parametrize("path", REFEXEC_IR, ids=[os.path.basename(f)[11:-4] for f in REFEXEC_IR])
def test_transport_reproduces_the_reference_executed_steps(path, lib_path):
    from mpas_seaice_b200 import irmesh, meshgen
    z = np.load(path)
    spec = ast.literal_eval(str(z["spec"]))
    mesh = getattr(meshgen, spec[0])(*spec[1:])
    irf = irmesh.ir_fields(mesh)
    rotate = bool(z["rotate"]) if "rotate" in z.files else False
    nqp = int(z["nqp"]) if "nqp" in z.files else 6
    geom = ir.init_geometry(mesh, irf, rotate=rotate)
    nC = mesh.nCells
    tracers = []
    for i in range(int(z["n_tracers"])):
        name, parent, vol = ast.literal_eval(str(z["meta_%d" % i]))
        tracers.append(ir.Tracer(name, z["in_%d" % i].copy(), parent, vol))
    u, v, dt, checks = z["in_uVelocity"], z["in_vVelocity"], float(z["dt"]), bool(z["checks"])
    ref, dev = clone(tracers), clone(tracers)
    solver = ir_host.IrTransport(mesh, irf, geom, tracers[0].array.shape[1], n_quad_points=nqp, rotate=rotate, lib_path=lib_path)
    try:
        solver.set_tracers(dev)
        if checks:
            solver.set_checks(conservation=1, monotonicity=1)
        for step in range(1, int(z["nsteps"]) + 1):
            # monotonicity_check = 1: the reference's own in-place extension of the bounds
            d = ir.run(mesh, irf, geom, ref, u, v, dt, n_quad_points=nqp, rotate=rotate, check=False,
                       conservation_check=int(checks), monotonicity_check=int(checks))
            rc = solver.run(dev, u, v, dt, check=False)
            for i, (x, y) in enumerate(zip(ref, dev)):
                want = z["out%d_%d" % (step, i)]
                assert np.array_equal(x.array[:nC], want[:nC]), ("oracle", x.name, step)
                assert np.array_equal(y.array[:nC], want[:nC]), ("kernels", x.name, step)
            if checks:
                aborts = [bool(b) for b in z["out%d_aborts" % step]]        # [.., update, conservation, monotonicity]
                assert d["error"] == (9 if aborts[-2] else (10 if aborts[-1] else 0))
                for i, t in enumerate(tracers):
                    nl = t.array.shape[2]
                    for key, mine in (("sumInit", d["sumInit"][i]), ("sumFinal", d["sumFinal"][i])):
                        want = z["out%d_%s_%d" % (step, key, i)].reshape(mine.shape)
                        assert np.array_equal(mine, want), (key, t.name)                # the oracle adds like the reference
                    si, sf = solver.conservation_sums(i, nl)
                    for mine, key in ((si, "sumInit"), (sf, "sumFinal")):
                        want = z["out%d_%s_%d" % (step, key, i)].reshape(mine.shape)
                        assert np.all(np.abs(mine - want) <= 1e-13 * np.abs(want).max() + 1e-300), (key, t.name)
                if aborts[-1]:      # the kernels' order-independent bounds are never looser than the reference's
                    assert rc == ir_host.IR_ERR_MONOTONICITY
                elif not aborts[-2]:
                    assert rc in (0, ir_host.IR_ERR_MONOTONICITY)
            else:
                assert rc == 0 and d["error"] == 0
        assert np.abs(ref[0].array[:nC] - tracers[0].array[:nC]).max() > 1e-6
    finally:
        solver.destroy()


REFEXEC_IR_INIT = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ir", "refexec_irinit_*.npz")))


@pytest.mark.parametrize("path", REFEXEC_IR_INIT, ids=[os.path.basename(f)[15:-4] for f in REFEXEC_IR_INIT])
def test_geometry_reproduces_the_reference_executed_init(path, lib_path):
    """seaice_init_advection_incremental_remap (incremental_remap.F:165-816) as the reference's source executes it:
    local frames, vertex coordinates in cell and edge frames, remap stencils, minimum edge lengths, the fourteen
    geometric cell averages -- reproduced bit for bit by the oracle and by ir_init_geometry."""
    from mpas_seaice_b200 import irmesh, meshgen
    z = np.load(path)
    assert "interpreting the reference's Fortran source" in str(z["provenance"]) and "get_geometry_incremental_remap" in str(z["provenance"])
    spec = ast.literal_eval(str(z["spec"]))
    mesh = getattr(meshgen, spec[0])(*spec[1:])
    assert np.array_equal(mesh.xCell, z["mesh_xCell"])
    irf = irmesh.ir_fields(mesh)
    rotate = bool(z["rotate"])
    nC = mesh.nCells
    for who, got in (("oracle", ir.init_geometry(mesh, irf, rotate=rotate)),
                     ("kernels", ir_host.init_geometry(mesh, irf, rotate=rotate, lib_path=lib_path))):
        for k in ("xVertexOnCell", "yVertexOnCell", "remapEdge", "cellsOnEdgeRemap", "edgesOnEdgeRemap", "xVertexOnEdge",
                  "yVertexOnEdge", "minLengthEdgesOnVertex"):
            n = got[k].shape[0] - 1
            assert np.array_equal(got[k][:n], z["out_" + k][:n]), (who, k)
        if mesh.on_a_sphere:
            assert np.array_equal(got["transGlobalToCell"][:nC], z["out_transGlobalToCell"][:nC]), who
        for name in ir.GEOM_NAMES:
            assert np.array_equal(got["geomAvg"][name][:nC], z["out_" + name + "AvgCell"][:nC]), (who, name)
    assert np.abs(z["out_xxAvgCell"][:nC]).max() > 0 and (z["out_remapEdge"] == 1).any()
    assert list(z["weightQuadPoint"]) == [1.09951743655321885e-01] * 3 + [2.23381589678011389e-01] * 3
