"""The vectorised numpy host helpers (variational_init.py, synthetic.py) against the oracle's loop-for-loop
restatement of the same reference routines: integer maps bit-exact, FP64 fields bit-exact where the
operation order is the reference's, else <= 2 ulp (documented per field)."""
import numpy as np
import pytest

import common
import oracle
from mpas_seaice_b200 import meshgen, synthetic, variational_init, workloads


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico3", "ico5"])
def test_init_static_matches_oracle(kind):
    mesh, var = common.mesh_case(kind)
    st = variational_init.init_static(mesh)
    assert np.array_equal(st["cellVerticesAtVertex"], var["cellVerticesAtVertex"])
    assert np.array_equal(st["interiorVertex"], var["interiorVertex"])
    assert np.array_equal(st["variationalDenominator"], var["variationalDenominator"])
    assert np.array_equal(st["xLocal"], var["xLocal"])
    assert np.array_equal(st["yLocal"], var["yLocal"])
    # tan(asin(z/R))/R: numpy and libm may differ in the last bit of tan / asin
    a, b = st["tanLatVertexRotatedOverRadius"], var["tanLatVertexRotatedOverRadius"]
    assert np.allclose(a, b, rtol=1e-14, atol=0.0)   # tan amplifies the asin ulp near the rotated poles


@pytest.mark.parametrize("kind", ["hex20", "ico3"])
def test_pre_subcycle_matches_oracle(kind):
    mesh, _ = common.mesh_case(kind)
    state = synthetic.sphere_state(mesh, "B") if mesh.on_a_sphere else synthetic.square_state(mesh)
    f, opts = synthetic.pre_subcycle(mesh, state, 3600.0)
    g = oracle.pre_subcycle(mesh, state, 3600.0)
    nV = mesh.nVertices
    for k in ("solveStress", "solveVelocity", "solveVelocityPrevious"):
        assert np.array_equal(f[k], g[k]), k
    vm = f["solveVelocity"][:nV] == 1
    assert vm.sum() > 0 and (~vm).sum() > 0
    for k in ("iceAreaVertex", "totalMassVertex", "uOceanVelocityVertex", "vOceanVelocityVertex", "airStressVertexU",
              "airStressVertexV", "totalMassVertexfVertex", "oceanStressU", "oceanStressV", "surfaceTiltForceU",
              "surfaceTiltForceV", "uVelocity", "vVelocity", "uVelocityInitial", "vVelocityInitial"):
        # consumers read these only where solveVelocity == 1
        assert np.array_equal(f[k][:nV][vm], g[k][:nV][vm]), k
    # exp() of numpy vs libm: last-bit differences allowed
    assert np.allclose(f["icePressure"], g["icePressure"], rtol=4e-16, atol=0.0)
    for k in ("stress11", "stress22", "stress12", "strain11", "replacementPressure", "stressDivergenceU"):
        assert np.array_equal(f[k], g[k]), k
    assert opts["elasticTimeStep"] == 30.0 and opts["dampingTimescale"] == 0.36 * 3600.0
    L = oracle.lib()
    assert opts["dampingTimescale"] == L.orc_damping_timescale(3600.0)
    dv = float(mesh.dvEdge[:-1].min())
    assert opts["numericalInertiaCoefficient"] == L.orc_numerical_inertia_coefficient(3600.0, dv)


def test_mesh_generators_are_valid_mpas_meshes():
    for mesh in (meshgen.planar_hex(12, 14, 1000.0), meshgen.planar_quad(9, 7, 500.0), meshgen.icosphere(2)):
        meshgen.check_mesh(mesh)
        nC, nV = mesh.nCells, mesh.nVertices
        assert mesh.areaCell[nC] == meshgen.JUNK_AREA and mesh.nEdgesOnCell[nC] == 0
        if mesh.on_a_sphere:
            assert nC == 10 * 4 ** 2 + 2 and nV == 20 * 4 ** 2 and (mesh.nEdgesOnCell[:nC] == 5).sum() == 12
            R = mesh.sphere_radius
            assert abs(mesh.areaCell[:nC].sum() / (4 * np.pi * R * R) - 1.0) < 1e-12
            assert abs(mesh.areaTriangle[:nV].sum() / (4 * np.pi * R * R) - 1.0) < 1e-12
        else:
            # kites tile the cells: interior vertices' triangles + boundary remainders = total cell area
            assert abs(mesh.kiteAreasOnVertex.sum() / mesh.areaCell[:nC].sum() - 1.0) < 1e-12


def test_workload_registry():
    w = workloads.build("square")
    mesh = w["mesh"]
    assert (mesh.nx, mesh.ny, mesh.dc) == (82, 94, 16000.0) and w["config_dt"] == 3600.0   # BASELINE configs[1]
    assert w["opts"]["elasticTimeStep"] == 30.0 and w["opts"]["n_elastic"] == 120
    nC_act, nV_act = workloads.active_counts(w)
    assert 0 < nC_act <= mesh.nCells and 0 < nV_act < mesh.nVertices
    assert workloads.SPHERES["qu7.5"] == (10, 120.0) and workloads.SPHERES["qu60"] == (7, 900.0)
    assert 10 * 4 ** 10 + 2 == 10485762
    with pytest.raises(ValueError):
        workloads.build("nope")


@pytest.mark.parametrize("kind", ["hex20", "quad40", "ico3"])
def test_weak_mesh_helper_geometry(kind):
    """weakmesh.weak_fields (synthetic stand-in for the mesh file's verticesOnEdge / edgesOnVertex and for
    seaice_normal_vectors): connectivity is consistent and the unit normals close every polygon / dual triangle."""
    from mpas_seaice_b200 import weakmesh
    mesh, _ = common.mesh_case(kind)
    w = weakmesh.weak_fields(mesh)
    nC, nV, nE, M, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges, mesh.vertexDegree
    voe, eov = w["verticesOnEdge"], w["edgesOnVertex"]
    assert np.all((voe[:nE] >= 1) & (voe[:nE] <= nV))
    # every edge of a cell joins two of that cell's vertices; both cells of an edge contain both vertices
    for c in (0, nC // 2, nC - 1):
        verts = set(mesh.verticesOnCell[c, :mesh.nEdgesOnCell[c]])
        for k in range(mesh.nEdgesOnCell[c]):
            e = mesh.edgesOnCell[c, k] - 1
            assert set(voe[e]) <= verts
    # an edge listed at a vertex has that vertex as one of its ends
    v = np.arange(nV)
    for s in range(D):
        e = eov[:nV, s]
        ok = e <= nE
        assert np.all((voe[e[ok] - 1, 0] == v[ok] + 1) | (voe[e[ok] - 1, 1] == v[ok] + 1))
    nvp, nvt = w["normalVectorPolygon"], w["normalVectorTriangle"]
    slot = np.arange(M)[None, :] < mesh.nEdgesOnCell[:nC, None]
    assert np.allclose(np.hypot(nvp[:nC, :, 0], nvp[:nC, :, 1])[slot], 1.0, rtol=1e-12)
    acc = np.zeros((nC, 2))
    for k in range(M):
        ok = slot[:, k]
        acc[ok] += nvp[:nC][ok, k, :] * mesh.dvEdge[mesh.edgesOnCell[:nC][ok, k] - 1][:, None]
    tol = 1e-12 if not mesh.on_a_sphere else 2e-2            # tangent-plane projection on the sphere
    assert np.abs(acc).max() <= tol * mesh.dvEdge[:nE].max()
    interior = np.all(mesh.cellsOnVertex[:nV] <= nC, axis=1)
    acc = np.zeros((nV, 2))
    for s in range(D):
        e = eov[:nV, s] - 1
        ok = e < nE
        acc[ok] += nvt[:nV][ok, s, :] * mesh.dcEdge[e[ok]][:, None]
    assert np.abs(acc[interior]).max() <= tol * mesh.dcEdge[:nE].max()


def test_interior_maps_against_the_reference_executed_init_boundary():
    """interiorVertex / interiorCell / interiorEdge as the reference's own init_boundary (src/shared/mpas_seaice_mesh.F:372-630)
    computes them -- its Fortran source interpreted by tests/golden/fortran_subset.py, fixtures
    tests/golden/options/refexec_boundary_*.npz -- against the host-side maps (variational_init.interior_vertex,
    ir_host.interior_edge) and the oracle's orc_interior_vertices.  Integer maps: bit-exact."""
    import ast
    import glob
    import os
    import oracle
    from mpas_seaice_b200 import ir_host, meshgen
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options", "refexec_boundary_*.npz")))
    assert len(files) == 4
    for f in files:
        z = np.load(f)
        assert "interior_vertices" in str(z["provenance"]) and "interior_edges" in str(z["provenance"])
        spec = ast.literal_eval(str(z["spec"]))
        mesh = getattr(meshgen, spec[0])(*spec[1:])
        assert np.array_equal(mesh.xCell, z["mesh_xCell"])
        nC, nV, nE, D = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.vertexDegree
        assert np.array_equal(variational_init.interior_vertex(mesh)[:nV], z["out_interiorVertex"][:nV]), f
        assert np.array_equal(ir_host.interior_edge(mesh)[:nE], z["out_interiorEdge"][:nE]), f
        mine = np.zeros(nV + 1, dtype=np.int32)
        oracle.lib().orc_interior_vertices(oracle._p(mine), nV, D, nC, oracle._p(mesh.cellsOnVertex))
        assert np.array_equal(mine[:nV], z["out_interiorVertex"][:nV]), f
        coc = mesh.cellsOnCell[:nC]
        used = np.arange(mesh.maxEdges)[None, :] < mesh.nEdgesOnCell[:nC, None]
        assert np.array_equal(np.all((coc <= nC) | ~used, axis=1).astype(np.int32), z["out_interiorCell"][:nC]), f


def test_locked_cells_against_the_reference_executed_mask():
    """dynamically_locked_cell_mask (src/shared/mpas_seaice_velocity_solver.F:402-467) after interior_vertices
    (mesh.F:423-488), both interpreted from the reference's source on a quadrilateral mesh with two culled columns
    (tests/golden/cpu/refexec_locked_cells.npz): the cells of the one-cell-wide channel between them have no interior
    vertex.  Host function and the oracle's restatement, bit-exact (an integer map of SURVEY section 8(c))."""
    import ast
    import os
    import oracle
    from mpas_seaice_b200 import meshgen
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cpu", "refexec_locked_cells.npz"))
    assert "dynamically_locked_cell_mask" in str(z["provenance"]) and "interior_vertices" in str(z["provenance"])
    spec = ast.literal_eval(str(z["spec"]))
    mesh = meshgen.Mesh(getattr(meshgen, spec[0])(*spec[1:]))
    nC, nV = mesh.nCells, mesh.nVertices
    for k in ("cellsOnVertex", "cellsOnCell"):
        a = mesh[k].copy()
        a[np.isin(a, z["culled"])] = nC + 1
        mesh[k] = a
    interior = variational_init.interior_vertex(mesh)
    assert np.array_equal(interior[:nV], z["out_interiorVertex"][:nV])
    ref = z["out_dynamicallyLockedCellsMask"][:nC]
    assert 0 < ref.sum() < nC
    assert np.array_equal(variational_init.dynamically_locked_cells_mask(mesh, interior)[:nC], ref)
    assert np.array_equal(oracle.dynamically_locked_cells_mask(mesh, interior)[:nC], ref)


def test_special_boundary_sources_against_the_reference_executed_init():
    """vertexBoundarySourceLocal / tracerBoundarySourceLocal as seaice_init_special_boundaries
    (src/shared/mpas_seaice_special_boundaries.F:60-250) computes them and the category tracers after
    seaice_set_special_boundaries_tracers (:415-485), both interpreted from the reference's source on a block whose local
    numbering is a permutation of the global IDs (tests/golden/cpu/refexec_special_boundaries_init.npz).  Integer maps
    bit-exact; the tracer copy is exact too (no arithmetic)."""
    import os
    import oracle
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cpu", "refexec_special_boundaries_init.npz"))
    for name in ("init_special_boundaries_velocity", "init_special_boundaries_tracers", "seaice_set_special_boundaries_tracers"):
        assert name in str(z["provenance"]), name
    for ids, btype, src, ref in (("indexToVertexID", "vertexBoundaryType", "vertexBoundarySource", "out_vertexBoundarySourceLocal"),
                                 ("indexToCellID", "tracerBoundaryType", "tracerBoundarySource", "out_tracerBoundarySourceLocal")):
        special = z[btype] != 0
        assert special.sum() >= 10
        assert not np.array_equal(z[ids][:-1], np.arange(1, len(z[ids])))        # local numbering differs from the global
        for f in (variational_init.boundary_source_local, oracle.boundary_source_local):
            got = f(z[ids], z[btype], z[src])
            assert np.array_equal(got[special], z[ref][special]), (ref, f.__module__)
            assert np.array_equal(z[ids][got[special] - 1], z[src][special])      # and it IS the inverse map
    names = ("iceAreaCategory", "iceVolumeCategory", "snowVolumeCategory")
    btype, loc = z["tracerBoundaryType"], z["out_tracerBoundarySourceLocal"]
    host = [z["in_" + k].copy() for k in names]
    orc = [z["in_" + k].copy() for k in names]
    variational_init.set_special_boundaries_tracers(btype, loc, *host)
    oracle.set_special_boundaries_tracers(btype, loc, *orc)
    for k, a, b in zip(names, host, orc):
        assert np.array_equal(a[:-1], z["out_" + k][:-1]), k
        assert np.array_equal(b[:-1], z["out_" + k][:-1]), k
        assert not np.array_equal(z["in_" + k], z["out_" + k])
    # the chained cells of the fixture: a SET cell whose source was emptied earlier in the loop comes out empty
    chained = [i for i in np.flatnonzero(btype[:-1] == 2) if btype[loc[i] - 1] == 1 and loc[i] - 1 < i]
    assert chained and all(np.all(z["out_iceAreaCategory"][i] == 0.0) for i in chained)
    with pytest.raises(ValueError):
        variational_init.boundary_source_local(z["indexToCellID"] + 5, btype, z["tracerBoundarySource"])
    with pytest.raises(ValueError):
        oracle.boundary_source_local(z["indexToCellID"] + 5, btype, z["tracerBoundarySource"])


def test_evp_parameters_against_the_reference_executed_seaice_init_evp():
    """constitutiveRelationType, dampingTimescale and numericalInertiaCoefficient as the reference's own seaice_init_evp
    (src/shared/mpas_seaice_velocity_solver_constitutive_relation.F:75-164) leaves them -- interpreted from its source,
    fixture tests/golden/options/refexec_init_evp.npz -- against the host-side synthetic.time_steps /
    numerical_inertia_coefficient every test's options come from, and the oracle's orc_* functions.  Bit for bit."""
    import ast
    import ctypes as C
    import os
    import oracle
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options", "refexec_init_evp.npz"))
    assert "seaice_init_evp" in str(z["provenance"])
    L = oracle.lib()
    L.orc_numerical_inertia_coefficient.restype = C.c_double
    L.orc_damping_timescale.restype = C.c_double
    rows = [ast.literal_eval(str(r)) for r in z["rows"]]
    assert len(rows) == 12
    for kind, cr, dt_dyn, nsub, cr_type, damping, inertia, dv_min in rows:
        assert cr_type == {"evp": oracle.EVP, "evp_revised": oracle.EVP_REVISED, "linear": oracle.LINEAR, "none": oracle.NONE}[cr]
        dt, dte, damp = synthetic.time_steps(dt_dyn, 1, nsub)
        assert damp == damping and synthetic.numerical_inertia_coefficient(dt, dv_min) == inertia
        assert L.orc_damping_timescale(C.c_double(dt_dyn)) == damping
        assert L.orc_numerical_inertia_coefficient(C.c_double(dt_dyn), C.c_double(dv_min)) == inertia


def test_square_test_case_state_against_the_reference_executed_routines():
    """The synthetic workload of BASELINE configs[1]: synthetic.square_state against init_square_test_case_state / _atmos /
    _ocean of the reference (src/shared/mpas_seaice_testing.F), interpreted from source
    (tests/golden/options/refexec_square_testcase.npz).  Inputs of the path, not results of it: held to round-off (numpy
    evaluates sin() in vector form, whose last bit may differ from libm's)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options", "refexec_square_testcase.npz"))
    assert "init_square_test_case_atmos" in str(z["provenance"]) and "init_square_test_case_state" in str(z["provenance"])
    m = meshgen.Mesh()
    m.xCell, m.yCell = z["in_x"], z["in_y"]
    st = synthetic.square_state(m)
    n = len(z["in_x"]) - 1
    for k in ("uAirVelocity", "vAirVelocity", "airDensity", "uOceanVelocity", "vOceanVelocity"):
        assert np.allclose(st[k][:n], z["out_" + k][:n], rtol=1e-14, atol=1e-15), k
    assert np.array_equal(st["iceAreaCell"][:n], z["out_iceAreaCategory"][:n, 0, 0])
    assert np.array_equal(st["iceVolumeCell"][:n], z["out_iceVolumeCategory"][:n, 0, 0])
    assert np.array_equal(st["snowVolumeCell"][:n], z["out_snowVolumeCategory"][:n, 0, 0])
    assert z["out_iceAreaCategory"].max() == 1.0 and np.abs(z["out_uAirVelocity"]).max() > 5.0
