"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerance: the north star allows 1e-10 relative max-norm after one dynamics step.  Because the
kernels are built with --fmad=false and keep the reference's operation order, the observed
difference against the non-FMA oracle is exactly zero; the tests assert BIT equality
(np.array_equal) on every compared field over owned entities with active masks, and fall back to
reporting the relative error in the assertion message.
"""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

TOL = 1e-10   # stated north-star tolerance (relative max-norm); observed: 0 (bit-exact)


def _compare(mesh, step, ref, out, fields_cell=common.COMPARE_CELL, fields_vertex=common.COMPARE_VERTEX):
    cm, vm = common.masks_for(mesh, step)
    worst = 0.0
    for k in fields_cell:
        err = common.rel_max_err(out[k], ref[k], cm)
        worst = max(worst, err)
        assert err <= TOL, f"{k}: rel max err {err:.3e}"
        assert np.array_equal(out[k][cm], ref[k][cm]), f"{k} not bit-exact (rel err {err:.3e})"
    for k in fields_vertex:
        err = common.rel_max_err(out[k], ref[k], vm)
        worst = max(worst, err)
        assert err <= TOL, f"{k}: rel max err {err:.3e}"
        assert np.array_equal(out[k][vm], ref[k][vm]), f"{k} not bit-exact (rel err {err:.3e})"
    return worst


@pytest.mark.parametrize("kind,nsub", [("hex20", 1), ("hex82", 1), ("hex82", 120), ("quad40", 120),
                                       ("ico3", 120), ("ico5", 120)])
def test_evp_subcycles_match_oracle(evp_lib, kind, nsub):
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    out = common.run_device(mesh, var, step, opts, nsub)
    _compare(mesh, step, ref, out)
    # inactive entities keep what the host gave (zero after init_subcycle_variables)
    cm, vm = common.masks_for(mesh, step)
    nV = mesh.nVertices
    assert np.all(out["uVelocity"][:nV][~vm[:nV]] == 0.0)


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico3", "ico5"])
def test_device_wachspress_precompute_bit_exact(evp_lib, kind):
    """evp_precompute_wachspress against the oracle's seaice_init_velocity_solver_wachspress
    (wachspress.F:46-161): all five basis arrays bit-identical, then a run from the device basis."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]))
    try:
        got = solver.fetch_basis()
        nC = mesh.nCells
        for k, a in got.items():
            assert np.array_equal(a[:nC], var[k][:nC]), k
        solver.update_step(step)
        solver.run_subcycles(5)
        out = solver.fetch()
    finally:
        solver.destroy()
    ref = common.run_oracle(mesh, var, step, opts, 5)
    _compare(mesh, step, ref, out)


@pytest.mark.parametrize("itype,order", [("dunavant", 1), ("dunavant", 4), ("dunavant", 7), ("trapezoidal", 3)])
def test_device_precompute_other_quadratures(evp_lib, itype, order):
    from mpas_seaice_b200 import host
    import oracle
    mesh, _ = common.mesh_case("hex20")
    var = oracle.init_variational(mesh, integration_type=itype, integration_order=order)
    step, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]), integration=(itype, order))
    try:
        got = solver.fetch_basis()
    finally:
        solver.destroy()
    for k, a in got.items():
        assert np.array_equal(a[:mesh.nCells], var[k][:mesh.nCells]), k


@pytest.mark.parametrize("cr,ocean", [("evp_revised", "quadratic"), ("evp", "linear"), ("linear", "quadratic"),
                                      ("none", "quadratic")])
@pytest.mark.parametrize("kind", ["hex20", "ico3"])
def test_namelist_options(evp_lib, kind, cr, ocean):
    """config_constitutive_relation_type x config_ocean_stress_type behind the boundary."""
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh, constitutive_relation_type=cr)
    opts = dict(opts, ocean_stress_type=ocean)
    if cr in ("linear", "none"):
        # operator-test style: a non-trivial velocity field that stays fixed (velocity_solver.F:2529-2541)
        nV = mesh.nVertices
        x = np.arange(nV + 1, dtype=np.float64)
        step["uVelocity"] = np.where(step["solveVelocity"] == 1, 0.1 * np.sin(0.37 * x), 0.0)
        step["vVelocity"] = np.where(step["solveVelocity"] == 1, 0.1 * np.cos(0.11 * x), 0.0)
        if cr == "none":
            step["stress11"][:] = np.where(step["solveStress"][:, None] == 1, 3.0, 0.0)
            step["stress12"][:] = np.where(step["solveStress"][:, None] == 1, -1.5, 0.0)
    nsub = 30
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    out = common.run_device(mesh, var, step, opts, nsub)
    fields_cell = common.COMPARE_CELL if cr != "none" else ("stress11", "stress22", "stress12", "strain11",
                                                            "strain22", "strain12")
    _compare(mesh, step, ref, out, fields_cell=fields_cell)
    if cr in ("linear", "none"):
        assert np.array_equal(out["uVelocity"], step["uVelocity"])


def test_no_ocean_stress(evp_lib):
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh, use_ocean_stress=False)
    ref = common.run_oracle(mesh, var, step, opts, 20)
    out = common.run_device(mesh, var, step, opts, 20)
    _compare(mesh, step, ref, out)
    assert np.all(out["oceanStressCoeff"] == 0.0)


@pytest.mark.parametrize("kind,state", [("ico5", "B"), ("hex82", "square")])
def test_partial_ice_cover_masks(evp_lib, kind, state):
    """State B (ice caps) / the square ramp: inactive cells and vertices next to active ones."""
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh, state_kind=state)
    nC, nV = mesh.nCells, mesh.nVertices
    assert 0 < (step["solveStress"][:nC] == 1).sum() < nC or state == "square"
    assert 0 < (step["solveVelocity"][:nV] == 1).sum() < nV
    ref = common.run_oracle(mesh, var, step, opts, 120)
    out = common.run_device(mesh, var, step, opts, 120)
    _compare(mesh, step, ref, out)
    off = step["solveStress"][:nC] != 1
    for k in ("stress11", "stress22", "stress12", "strain11", "replacementPressure"):
        assert np.all(out[k][:nC][off] == 0.0), k


def test_pwl_basis_dense_gradients(evp_lib):
    """config_variational_basis = 'pwl': gradients are dense (pwl.F:259-274), uploaded from the host."""
    mesh, var = common.mesh_case("ico3", basis="pwl")
    step, opts = common.step_case(mesh)
    ref = common.run_oracle(mesh, var, step, opts, 60)
    out = common.run_device(mesh, var, step, opts, 60)
    _compare(mesh, step, ref, out)


def test_alternate_denominator(evp_lib):
    mesh, var = common.mesh_case("ico3", denominator="alternate")
    step, opts = common.step_case(mesh)
    ref = common.run_oracle(mesh, var, step, opts, 60)
    out = common.run_device(mesh, var, step, opts, 60)
    _compare(mesh, step, ref, out)


def test_state_carried_across_dynamics_steps(evp_lib):
    """Two dynamics steps: u, v, stresses and solveVelocityPrevious carried, the second
    step starts from non-zero stress; CUDA-graph replay (second call reuses the instantiated graph)."""
    from mpas_seaice_b200 import host, synthetic
    mesh, var = common.mesh_case("ico3")
    state = synthetic.sphere_state(mesh, "A")
    step, opts = synthetic.pre_subcycle(mesh, state, 3600.0)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        prev_dev, prev_ref = None, None
        for it in range(2):
            s_dev, _ = synthetic.pre_subcycle(mesh, state, 3600.0, prev=prev_dev)
            s_ref, _ = synthetic.pre_subcycle(mesh, state, 3600.0, prev=prev_ref)
            solver.update_step(s_dev)
            solver.run_subcycles(120)
            out = solver.fetch()
            ref = common.run_oracle(mesh, var, s_ref, opts, 120)
            _compare(mesh, s_ref, ref, out)
            prev_dev = dict(out, solveVelocityPrevious=s_dev["solveVelocityPrevious"])
            prev_ref = dict(ref, solveVelocityPrevious=s_ref["solveVelocityPrevious"])
        assert np.abs(s_ref["stress11"]).max() > 0
    finally:
        solver.destroy()


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico4"])
def test_graph_and_stream_paths_agree(evp_lib, kind, monkeypatch):
    """Three ways to run the same subcycles: the persistent whole-loop kernel (one cooperative launch, the default for
    meshes whose tiles are all resident at once), the CUDA graph of cell / vertex kernel nodes, plain stream launches."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    outs = []
    for mode, use_graph, persistent, launches in (("persistent", 1, "1", 1), ("graph", 1, "0", 20), ("stream", 0, "1", 20)):
        monkeypatch.setenv("EVP_B200_PERSISTENT", persistent)
        solver = host.EvpSolver(mesh, var, opts)
        try:
            solver.set_use_graph(use_graph)
            solver.update_step(step)
            solver.run_subcycles(7)
            solver.run_subcycles(3)          # a second call with a different count re-instantiates
            outs.append(solver.fetch())
            assert solver.launch_count(10) == launches, mode
        finally:
            solver.destroy()
    for k in common.COMPARE_CELL + common.COMPARE_VERTEX:
        assert np.array_equal(outs[0][k], outs[1][k]), k
        assert np.array_equal(outs[0][k], outs[2][k]), k
    ref = common.run_oracle(mesh, var, step, opts, 10)
    # strain / divergence diagnostics are those of the LAST subcycle in all of them
    _compare(mesh, step, ref, outs[0])


@pytest.mark.parametrize("cr", ["evp", "evp_revised"])
def test_persistent_kernel_with_partial_cover(evp_lib, cr, monkeypatch):
    """The whole-loop kernel on the polar-cap state (unsolved cells and vertices inside resident tiles) and with the
    revised EVP relation, 120 subcycles, against the oracle."""
    from mpas_seaice_b200 import host
    monkeypatch.setenv("EVP_B200_PERSISTENT", "1")
    mesh, var = common.mesh_case("ico5")
    step, opts = common.step_case(mesh, state_kind="B", constitutive_relation_type=cr)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        solver.update_step(step)
        solver.run_subcycles(120)
        out = solver.fetch()
        assert solver.launch_count(120) == 1
    finally:
        solver.destroy()
    _compare(mesh, step, common.run_oracle(mesh, var, step, opts, 120), out)


def test_pinned_host_path(evp_lib):
    """EVP_FLAG_PIN_HOST (cudaHostRegister of the caller's arrays) gives the same bits as the bounce path."""
    mesh, var = common.mesh_case("ico5")
    step, opts = common.step_case(mesh)
    a = common.run_device(mesh, var, step, opts, 10)
    b = common.run_device(mesh, var, step, opts, 10, pin_host=True)
    for k in common.COMPARE_CELL + common.COMPARE_VERTEX:
        assert np.array_equal(a[k], b[k]), k


def _special_boundary_case():
    """1D_velocity_hex-like set-up (testing_and_setup/testcases/square/1D_velocity_hex): periodic in x
    (type 1), reversed copies (type 2) and zero-velocity vertices (type 3), with a chain
    (a boundary vertex whose source is itself a boundary vertex)."""
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh)
    nV = mesh.nVertices
    vbt = np.zeros(nV + 1, dtype=np.int32)
    src = np.zeros(nV + 1, dtype=np.int32)
    active = np.nonzero(step["solveVelocity"][:nV] == 1)[0]
    inactive = np.nonzero(step["solveVelocity"][:nV] != 1)[0]
    rng = np.random.default_rng(3)
    pick = rng.choice(inactive, size=min(60, inactive.size), replace=False)
    for i, v in enumerate(pick):
        t = 1 + (i % 3)
        vbt[v] = t
        if t in (1, 2):
            src[v] = int(rng.choice(active)) + 1
    # chains: earlier and later boundary vertices as sources
    chain = pick[vbt[pick] == 1]
    if chain.size >= 4:
        src[chain[0]] = chain[3] + 1     # source updated LATER in the sequential loop (old value seen)
        src[chain[2]] = chain[1] + 1     # source updated EARLIER (new value seen)
    return mesh, var, step, opts, vbt, src


def test_special_boundaries_velocity(evp_lib):
    """seaice_set_special_boundaries_velocity (special_boundaries.F:253-331) inside the captured loop."""
    mesh, var, step, opts, vbt, src = _special_boundary_case()
    opts = dict(opts, use_special_boundaries_velocity=True)
    ostep = common.clone_step(step)
    ostep["vertexBoundaryType"], ostep["vertexBoundarySourceLocal"] = vbt, src
    ref = common.run_oracle(mesh, var, ostep, opts, 25)
    out = common.run_device(mesh, var, step, opts, 25, special_boundaries=(vbt, src))
    _compare(mesh, step, ref, out)
    nV = mesh.nVertices
    b = vbt[:nV] != 0
    assert np.array_equal(out["uVelocity"][:nV][b], ref["uVelocity"][:nV][b])
    assert np.array_equal(out["vVelocity"][:nV][b], ref["vVelocity"][:nV][b])
    assert np.abs(ref["uVelocity"][:nV][b]).max() > 0


def test_set_masks(evp_lib):
    """seaice_set_special_boundaries_velocity_masks (special_boundaries.F:345-401): masks replaced by
    externally supplied arrays."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh)
    nC, nV = mesh.nCells, mesh.nVertices
    ss = step["solveStress"].copy()
    sv = step["solveVelocity"].copy()
    sv[:nV:7] = 0
    ss[:nC:5] = 0
    solver = host.EvpSolver(mesh, var, opts)
    try:
        solver.update_step(step)
        solver.set_masks(ss, sv)
        solver.run_subcycles(15)
        out = solver.fetch()
    finally:
        solver.destroy()
    ostep = common.clone_step(step)
    ostep["solveStress"], ostep["solveVelocity"] = ss, sv
    ref = common.run_oracle(mesh, var, ostep, opts, 15)
    _compare(mesh, ostep, ref, out)


def test_call_order_errors(evp_lib):
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("hex20")
    step, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        with pytest.raises(host.EvpError, match="update_step"):
            solver.run_subcycles(1)
        with pytest.raises(host.EvpError):
            solver.last_run_ms()
        bad = dict(step)
        bad["icePressure"] = None
        with pytest.raises(host.EvpError, match="NULL"):
            solver.update_step(bad)
        with pytest.raises(host.EvpError, match="special boundaries"):
            solver.set_options(dict(opts, use_special_boundaries_velocity=True))
    finally:
        solver.destroy()


def test_linearity_of_the_stress_divergence_at_full_size(evp_lib):
    """Size-independent property at a BASELINE-sized mesh the oracle would need minutes for (QU60,
    163 842 cells): with the linear constitutive relation the strain -> stress -> divergence chain is
    linear in (u, v): D(a*u1 + u2) == a*D(u1) + D(u2) to round-off, and a constant velocity on the
    rotated sphere has zero strain where the metric term vanishes is NOT assumed -- only linearity."""
    from mpas_seaice_b200 import host, workloads
    w = workloads.build("qu60")
    mesh, static, step, opts = w["mesh"], w["static"], w["step"], w["opts"]
    opts = dict(opts, constitutive_relation_type="linear")
    nV = mesh.nVertices
    lat, lon = mesh.latVertex, mesh.lonVertex
    u1, v1 = np.cos(lat) * np.sin(3 * lon), np.sin(2 * lat) * np.cos(lon)
    u2, v2 = np.sin(lat) ** 2 * np.cos(2 * lon), np.cos(lat) * np.sin(5 * lon)
    solver = host.EvpSolver(mesh, static, opts, local_coords=(static["xLocal"], static["yLocal"]))
    res = []
    try:
        for (u, v) in ((u1, v1), (u2, v2), (2.5 * u1 + u2, 2.5 * v1 + v2)):
            s = dict(step)
            s["uVelocity"], s["vVelocity"] = np.ascontiguousarray(u), np.ascontiguousarray(v)
            solver.update_step(s)
            solver.run_subcycles(1)
            res.append(solver.fetch(names=("stressDivergenceU", "stressDivergenceV", "strain12")))
    finally:
        solver.destroy()
    for k in ("stressDivergenceU", "stressDivergenceV", "strain12"):
        lhs = res[2][k]
        rhs = 2.5 * res[0][k] + res[1][k]
        scale = np.abs(rhs).max()
        assert scale > 0
        assert np.abs(lhs - rhs).max() <= 1e-12 * scale, k


@pytest.mark.parametrize("kind", ["hex20", "ico4", "quad40"])
def test_average_variational_strain(evp_lib, kind):
    """config_average_variational_strain: seaice_average_strains_on_vertex (variational.F:684-763) between the
    strain and the stress update -- the fused cell kernel splits into strain / vertex average / stress."""
    from mpas_seaice_b200 import host, variational_init
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    opts = dict(opts, average_variational_strain=True)
    nsub = 30
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    plain = common.run_oracle(mesh, var, step, dict(opts, average_variational_strain=False), nsub)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        solver.update_step(step)
        with pytest.raises(host.EvpError, match="evp_set_mesh_ext"):
            solver.run_subcycles(nsub)
        solver.set_mesh_ext(mesh, variational_init.interior_vertex(mesh))
        solver.run_subcycles(nsub)
        out = solver.fetch()
        assert solver.launch_count(nsub) == 4 * nsub
    finally:
        solver.destroy()
    _compare(mesh, step, ref, out)
    cm, _ = common.masks_for(mesh, step)
    assert not np.array_equal(ref["stress11"][cm], plain["stress11"][cm])      # the option does change the answer


def test_split_subcycle_counts_at_full_size(evp_lib):
    """Size-independent property at QU60 (BASELINE configs[3], 163 842 cells): 120 subcycles in one call, in two
    calls of 60, and without graph replay give the same bits for every carried field (u, v, stresses), because
    nothing but those fields is carried between subcycles."""
    from mpas_seaice_b200 import host, workloads
    w = workloads.build("qu60")
    mesh, static, step, opts = w["mesh"], w["static"], w["step"], w["opts"]
    names = ("uVelocity", "vVelocity", "stress11", "stress22", "stress12", "strain11", "stressDivergenceU")
    solver = host.EvpSolver(mesh, static, opts, local_coords=(static["xLocal"], static["yLocal"]))
    res = []
    try:
        for plan, graph in (((120,), 1), ((60, 60), 1), ((119, 1), 0)):
            solver.set_use_graph(graph)
            solver.update_step(step)
            for n in plan:
                solver.run_subcycles(n)
            res.append(solver.fetch(names=names))
    finally:
        solver.destroy()
    assert np.abs(res[0]["uVelocity"]).max() > 0
    for other in res[1:]:
        for k in names:
            assert np.array_equal(res[0][k], other[k]), k


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico4"])
def test_device_pwl_precompute_bit_exact(evp_lib, kind):
    """evp_precompute_pwl against the oracle's seaice_init_velocity_solver_pwl (pwl.F:44-373, LU solves of
    numerics.F:44-212): all five basis arrays bit-identical, then a run from the device basis (dense gradients)."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case(kind, basis="pwl")
    step, opts = common.step_case(mesh)
    solver = host.EvpSolver(mesh, var, opts, local_coords=(var["xLocal"], var["yLocal"]), basis="pwl")
    try:
        got = solver.fetch_basis()
        nC = mesh.nCells
        for k, a in got.items():
            assert np.array_equal(a[:nC], var[k][:nC]), k
        solver.update_step(step)
        solver.run_subcycles(20)
        out = solver.fetch()
    finally:
        solver.destroy()
    ref = common.run_oracle(mesh, var, step, opts, 20)
    _compare(mesh, step, ref, out)


def _pad_max_edges(mesh, var, step, new_m):
    """The same mesh stored with a larger maxEdges (MPAS meshes with a few 7-sided cells have maxEdges = 7):
    every (maxEdges, ...) array gets extra, unused columns."""
    from mpas_seaice_b200 import meshgen
    M = mesh.maxEdges
    nC = mesh.nCells

    def pad_last(a, fill):
        out = np.full(a.shape[:-1] + (new_m,), fill, dtype=a.dtype)
        out[..., :M] = a
        return np.ascontiguousarray(out)

    m2 = meshgen.Mesh(mesh)
    m2["maxEdges"] = new_m
    for k in ("verticesOnCell", "edgesOnCell", "cellsOnCell"):
        fill = {"verticesOnCell": mesh.nVertices + 1, "edgesOnCell": mesh.nEdges + 1, "cellsOnCell": nC + 1}[k]
        m2[k] = pad_last(mesh[k], fill)
    v2 = dict(var)
    for k in ("basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric"):
        a = np.zeros((nC + 1, new_m, new_m))
        a[:, :M, :M] = var[k]
        v2[k] = a
    for k in ("xLocal", "yLocal"):
        if k in var:
            v2[k] = pad_last(var[k], 0.0)
    s2 = dict(step)
    for k in ("stress11", "stress22", "stress12", "strain11", "strain22", "strain12", "replacementPressure"):
        s2[k] = pad_last(step[k], 0.0)
    return m2, v2, s2


@pytest.mark.parametrize("new_m", [7, 8])
def test_host_max_edges_larger_than_any_cell(evp_lib, new_m):
    """maxEdges = 7 (handled by the 8-slot instantiation) and 8 on a mesh whose cells have 5-6 vertices:
    the host layout (Mh) and the device layout (M) differ."""
    mesh, var = common.mesh_case("ico3")
    step, opts = common.step_case(mesh)
    m2, v2, s2 = _pad_max_edges(mesh, var, step, new_m)
    ref = common.run_oracle(m2, v2, s2, opts, 30)
    out = common.run_device(m2, v2, s2, opts, 30)
    _compare(m2, s2, ref, out)
    plain = common.run_oracle(mesh, var, step, opts, 30)
    assert np.array_equal(ref["uVelocity"], plain["uVelocity"])            # padding changes nothing
    # device Wachspress precompute with the padded local coordinates
    from mpas_seaice_b200 import host
    solver = host.EvpSolver(m2, v2, opts, local_coords=(v2["xLocal"], v2["yLocal"]))
    try:
        got = solver.fetch_basis()
    finally:
        solver.destroy()
    for k, a in got.items():
        assert np.array_equal(a[:mesh.nCells], v2[k][:mesh.nCells]), k


def test_no_ice_anywhere_and_empty_block(evp_lib):
    """Edge cases: every mask off (ice-free ocean: nothing is read but the masks, everything stays zero), and a
    block without any cell or vertex (a rank whose partition is empty)."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("ico3")
    step, opts = common.step_case(mesh)
    s = common.clone_step(step)
    s["solveStress"][:] = 0
    s["solveVelocity"][:] = 0
    for k in ("uVelocity", "vVelocity"):
        s[k][:] = 0.0
    out = common.run_device(mesh, var, s, opts, 10)
    for k in ("uVelocity", "vVelocity", "stress11", "stress12", "strain11", "replacementPressure", "stressDivergenceU"):
        assert not out[k].any(), k
    from mpas_seaice_b200 import meshgen
    empty = meshgen.Mesh(nCells=0, nVertices=0, maxEdges=6, vertexDegree=3,
                         nEdgesOnCell=np.zeros(1, dtype=np.int32), verticesOnCell=np.ones((1, 6), dtype=np.int32),
                         cellsOnVertex=np.ones((1, 3), dtype=np.int32))
    evar = dict(cellVerticesAtVertex=np.zeros((1, 3), dtype=np.int32), tanLatVertexRotatedOverRadius=np.zeros(1),
                variationalDenominator=np.zeros(1))
    for k in ("basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric"):
        evar[k] = np.zeros((1, 6, 6))
    estep = {k: (np.zeros((1, 6)) if v.ndim == 2 else np.zeros(1, dtype=v.dtype)) for k, v in step.items()
             if isinstance(v, np.ndarray)}
    solver = host.EvpSolver(empty, evar, opts)
    try:
        solver.update_step(estep)
        solver.run_subcycles(5)
        assert solver.launch_count(5) == 0
        res = solver.fetch()
    finally:
        solver.destroy()
    assert res["uVelocity"].shape == (1,)


def test_set_options_between_steps(evp_lib):
    """evp_set_options every dynamics step (what the Fortran shim does): unchanged options keep the instantiated
    graph, changed scalars (config_dt changed -> elasticTimeStep, dampingTimescale) rebuild it -- the time steps are
    kernel arguments baked into the graph nodes."""
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case("ico3")
    step, opts = common.step_case(mesh)
    opts2 = dict(opts, elasticTimeStep=opts["elasticTimeStep"] / 2.0, dynamicsTimeStep=opts["dynamicsTimeStep"] / 2.0,
                 dampingTimescale=opts["dampingTimescale"] / 2.0)
    solver = host.EvpSolver(mesh, var, opts)
    try:
        outs = []
        for o in (opts, opts, opts2, opts):
            solver.set_options(o)
            solver.update_step(step)
            solver.run_subcycles(40)
            outs.append(solver.fetch())
    finally:
        solver.destroy()
    ref1 = common.run_oracle(mesh, var, step, opts, 40)
    ref2 = common.run_oracle(mesh, var, step, opts2, 40)
    _compare(mesh, step, ref1, outs[0])
    _compare(mesh, step, ref1, outs[1])
    _compare(mesh, step, ref2, outs[2])
    _compare(mesh, step, ref1, outs[3])
    cm, vm = common.masks_for(mesh, step)
    assert not np.array_equal(ref1["uVelocity"][vm], ref2["uVelocity"][vm])
