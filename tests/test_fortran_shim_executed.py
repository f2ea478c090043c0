"""The ISO_C_BINDING shims under fortran/ EXECUTED: tests/golden/fortran_subset.py interprets the shim's Fortran source
and tests/golden/fortran_cbridge.py turns its bind(C) interface blocks into real calls of the shared library, marshalled
from the Fortran declarations (VALUE / by reference, type(c_ptr), bind(C) types as structs).  No Fortran compiler exists
in this image; this is the closest thing to running the drop-in: the host side is the shim's own statements, the pools
are the harness's arrays, the library is the product's.

* fortran/seaice_ir_b200.F90 against the host emulation of libir_b200.so here (CPU) and against the CUDA library on the
  GPU: seaice_ir_b200_create -> seaice_ir_b200_step (the tracer linked list flattened by the shim, the namelist checks
  passed on, the conservation sums stored where check_tracer_conservation reads them) -> seaice_upwind_b200_step ->
  seaice_ir_b200_destroy; results equal to the oracle's, bit for bit.
* fortran/seaice_evp_b200.F90 against libevp_b200.so on the GPU: seaice_evp_b200_create / _update / _subcycle /
  _destroy from the pools of a synthetic step; results equal to the oracle's, bit for bit.
"""
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import fortran_subset as F  # noqa: E402
from fortran_cbridge import CBridge  # noqa: E402

import common  # noqa: E402
from oracle import ir, upwind  # noqa: E402
from mpas_seaice_b200 import ir_host, variational_init  # noqa: E402
from test_oracle_ir import case, smooth_divergent_velocity, _random_state  # noqa: E402
from test_ir_parity import _emulation_library, clone  # noqa: E402


def _block_domain():
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None, localblockid=0)
    return block, types.SimpleNamespace(blocklist=block, configs="configs")


def _tracer_list(tracers, work, nK):
    """the reference's tracer_type linked list (incremental_remap_tracers.F:26-110) over the arrays of ``work``"""
    objs = []
    for i, t in enumerate(tracers):
        nl = t.array.shape[2]
        o = types.SimpleNamespace(tracername=t.name, ndims=2 if nl == 1 else 3, parent=None, next=None, nparents=0,
                                  array2d=F.FArray(work[i][:, :, 0]) if nl == 1 else None,
                                  array3d=F.FArray(work[i]) if nl > 1 else None)
        if nl == 1:
            o.globalsuminit2d, o.globalsumfinal2d = F.FArray(np.zeros(nK)), F.FArray(np.zeros(nK))
        else:
            o.globalsuminit3d, o.globalsumfinal3d = F.FArray(np.zeros((nK, nl))), F.FArray(np.zeros((nK, nl)))
        objs.append(o)
    for i, t in enumerate(tracers):
        if t.parent is not None:
            objs[i].parent, objs[i].nparents = objs[t.parent], objs[t.parent].nparents + 1
        if i + 1 < len(objs):
            objs[i].next = objs[i + 1]
    return objs


def _run_ir_shim(lib_path, kind="hex16", checks=False):
    mesh, irf, geom = case(kind)
    nC, nV, nE, M = mesh.nCells, mesh.nVertices, mesh.nEdges, mesh.maxEdges
    tracers = _random_state(mesh, np.random.default_rng(21))
    nK = tracers[0].array.shape[1]
    u, v = smooth_divergent_velocity(mesh, geom)
    I = F.Interpreter()
    I.load(os.path.join(ROOT, "fortran", "seaice_ir_b200.F90"))
    I.resolve_constants()
    I.noop |= {"seaice_set_tracer_array_pointers", "mpas_log_write"}
    I.globals["mpas_log_crit"] = 3
    bridge = CBridge(I, C.CDLL(lib_path))
    work = [t.array.copy() for t in tracers]
    objs = _tracer_list(tracers, work, nK)
    pool = {k: mesh[k] for k in list(mesh.keys()) if isinstance(mesh[k], np.ndarray)}
    pool.update(verticesOnEdge=irf["verticesOnEdge"], coeffs_reconstruct=irf["coeffs_reconstruct"], uVelocity=u, vVelocity=v)
    for k in ("transGlobalToCell", "xVertexOnCell", "yVertexOnCell", "xVertexOnEdge", "yVertexOnEdge", "remapEdge", "cellsOnEdgeRemap",
              "edgesOnEdgeRemap"):
        pool[k] = geom[k]
    for n in ir.GEOM_NAMES:
        pool[n + "AvgCell"] = geom["geomAvg"][n]
    # the upwind step's pools: (1, nCategories, nCells+1) tracers, interiorEdge, normalVectorEdge
    iv = variational_init.interior_vertex(mesh)
    nve = upwind.normal_vectors(mesh, irf, iv, rotate=True, remove_metric_terms=False, triangles=False)["normalVectorPolygon"]
    pool.update(interiorEdge=ir_host.interior_edge(mesh), normalVectorEdge=nve)
    byname = {t.name: work[i] for i, t in enumerate(tracers)}
    for k, a in pool.items():
        I.pool[k] = F.FArray(a)
    for name in ("iceAreaCategory", "surfaceTemperature", "iceVolumeCategory", "snowVolumeCategory"):
        I.pool[("tracers", name, 1)] = F.FArray(byname[name])
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nEdges=nE, maxEdges=M, vertexDegree=mesh.vertexDegree, nCategories=nK,
                  nQuadPoints=6, on_a_sphere=bool(mesh.on_a_sphere), config_rotate_cartesian_grid=False,
                  config_conservation_check=bool(checks), config_monotonicity_check=bool(checks))
    block, domain = _block_domain()
    I.call("seaice_ir_b200_create", block)
    return I, bridge, mesh, irf, geom, tracers, work, objs, u, v, block, domain, nve


def _check_ir_shim(lib_path):
    I, bridge, mesh, irf, geom, tracers, work, objs, u, v, block, domain, nve = _run_ir_shim(lib_path, checks=True)
    nC = mesh.nCells
    ref = clone(tracers)
    for step in range(2):
        d = ir.run(mesh, irf, geom, ref, u, v, 3600.0, conservation_check=2)
        I.call("seaice_ir_b200_step", domain, block, 3600.0, objs[0])
        for i, t in enumerate(ref):
            assert np.array_equal(work[i][:nC], t.array[:nC]), (t.name, step)
            # mode 2 of the conservation check: this block's sums, stored where check_tracer_conservation (:8126) reads them
            o = objs[i]
            mine = (o.globalsuminit2d if o.ndims == 2 else o.globalsuminit3d).a
            assert np.all(np.abs(mine.reshape(d["sumInit"][i].shape) - d["sumInit"][i]) <= 1e-13 * np.abs(d["sumInit"][i]).max())
            mine = (o.globalsumfinal2d if o.ndims == 2 else o.globalsumfinal3d).a
            assert np.all(np.abs(mine.reshape(d["sumFinal"][i].shape) - d["sumFinal"][i]) <= 1e-13 * np.abs(d["sumFinal"][i]).max())
    assert np.abs(work[0][:nC] - tracers[0].array[:nC]).max() > 1e-6
    assert bridge.calls[:5] == ["ir_create", "ir_set_tracers", "ir_set_checks", "ir_run", "ir_fetch_conservation_sums"]
    assert bridge.calls.count("ir_set_tracers") == 1                      # the table is declared once, then reused
    # the upwind option through the shim: the reference's table, in place on time level 1
    names = ("iceAreaCategory", "surfaceTemperature", "iceVolumeCategory", "snowVolumeCategory")
    idx = {t.name: i for i, t in enumerate(tracers)}
    parents = {"iceAreaCategory": None, "surfaceTemperature": 0, "iceVolumeCategory": 1, "snowVolumeCategory": 2}
    uref = [upwind.Var(n, work[idx[n]][:, :, 0].copy(), parents[n], n.endswith("VolumeCategory")) for n in names]
    upwind.run(mesh, irf["verticesOnEdge"], ir_host.interior_edge(mesh), nve, uref, u, v, 3600.0)
    I.call("seaice_upwind_b200_step", block, 3600.0)
    for x in uref:
        assert np.array_equal(work[idx[x.name]][:, :, 0], x.array), x.name
    assert "ir_set_upwind_mesh" in bridge.calls and "ir_run_upwind" in bridge.calls
    I.call("seaice_ir_b200_destroy")
    assert I.globals["irhandle"] is None and bridge.calls[-1] == "ir_destroy"


def test_ir_shim_executed_against_the_emulated_library():
    _check_ir_shim(_emulation_library())


@pytest.mark.gpu
def test_ir_shim_executed_against_the_cuda_library():
    _check_ir_shim(ir_host.LIB_PATH)


# ----------------------------------------------------------------------------------------------------------------------
# fortran/seaice_evp_b200.F90: seaice_evp_b200_create / _update / _subcycle / _destroy executed against libevp_b200.so
# ----------------------------------------------------------------------------------------------------------------------

def _evp_shim(kind="hex20", nsub=20, cr="evp", lib_path=None):
    from mpas_seaice_b200 import host
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh, constitutive_relation_type=cr)
    work = common.clone_step(step)
    I = F.Interpreter()
    I.load(os.path.join(ROOT, "fortran", "seaice_evp_b200.F90"))
    I.resolve_constants()
    I.noop |= {"seaice_set_special_boundaries_velocity_masks"}
    I.globals["mpas_log_crit"] = 3
    bridge = CBridge(I, C.CDLL(lib_path or host.LIB_PATH))
    bridge.log = []

    def c_f_pointer(interp, fr, args):            # call c_f_pointer(cptr, fptr, [n]): the C string as a character array
        addr, n = interp.ev(args[0][1], fr), int(interp.ev(args[2][1], fr).a[0])
        text = (C.string_at(addr) if addr else b"").decode()[:n - 1]
        fr.bind(args[1][1][1], F.FArray(np.array(list(text) + ["\0"] * (n - len(text)), dtype=object)))

    def log_write(interp, fr, args):
        bridge.log.append(interp.ev(args[0][1], fr))

    I.hooks.update(c_f_pointer=c_f_pointer, mpas_log_write=log_write)
    nC, nV = mesh.nCells, mesh.nVertices
    for k in ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex"):
        I.pool[k] = F.FArray(mesh[k])
    for k in ("cellVerticesAtVertex", "basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric",
              "tanLatVertexRotatedOverRadius", "variationalDenominator"):
        I.pool[k] = F.FArray(var[k])
    for k, a in work.items():
        if isinstance(a, np.ndarray):
            I.pool[k] = F.FArray(a)
    I.pool.update(nCells=nC, nCellsSolve=nC, nVertices=nV, nVerticesSolve=nV, maxEdges=mesh.maxEdges, vertexDegree=mesh.vertexDegree,
                  elasticTimeStep=float(opts["elasticTimeStep"]), dynamicsTimeStep=float(opts["dynamicsTimeStep"]),
                  config_ocean_stress_type="quadratic", config_use_ocean_stress=True, config_use_special_boundaries_velocity=False,
                  config_use_special_boundaries_velocity_masks=False, config_elastic_subcycle_number=int(nsub),
                  config_average_variational_strain=False, config_strain_scheme="variational",
                  config_stress_divergence_scheme="variational", pkgVariationalActive=True)
    # module variables of seaice_velocity_solver_constitutive_relation the shim uses (constitutive_relation.F:29-59)
    I.globals.update(constitutiverelationtype={"evp": 1, "evp_revised": 2, "linear": 3}[cr],
                     dampingtimescale=float(opts["dampingTimescale"]),
                     numericalinertiacoefficient=float(opts.get("numericalInertiaCoefficient", 0.0)))
    block = types.SimpleNamespace(structs="structs", configs="configs", dimensions="dimensions", next=None)
    domain = types.SimpleNamespace(blocklist=block, configs="configs", packages="packages")
    return I, bridge, domain, mesh, var, step, opts, work


def test_evp_shim_reaches_the_library_with_a_valid_descriptor():
    """Without a device evp_create must fail with EVP_ERR_CUDA (2) -- AFTER it has accepted the mesh descriptor and the
    options the shim built (an argument error would be 1): the bind(C) types and the interface of evp_create as the shim
    declares them are what the library expects."""
    I, bridge, domain, *_ = _evp_shim()
    assert I.call("seaice_evp_b200_supported", domain).vars["supported"] is True
    I.call("seaice_evp_b200_create", domain)
    name, rc = bridge.returns[0]
    assert name == "evp_create"
    import subprocess
    has_gpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0 if os.path.exists("/usr/bin/nvidia-smi") else False
    assert rc == (0 if has_gpu else 2), rc
    if rc == 0:
        I.call("seaice_evp_b200_destroy", 0)
    else:       # the shim's own error path: evp_last_error_string() copied character by character into the MPAS log line
        assert len(bridge.log) == 1 and bridge.log[0].startswith("libevp_b200: evp_create: ") and "cuda" in bridge.log[0].lower()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,cr", [("hex20", "evp"), ("ico3", "evp_revised")])
def test_evp_shim_executed_against_the_cuda_library(evp_lib, kind, cr):
    """The shim's own create -> update -> subcycle (evp_set_options, evp_update_step, evp_run_subcycles, evp_fetch into the
    pool arrays) -> destroy, from the pools of a synthetic step; the pool arrays afterwards hold the oracle's results."""
    nsub = 20
    I, bridge, domain, mesh, var, step, opts, work = _evp_shim(kind, nsub, cr, lib_path=evp_lib._name)   # (the CUDA build, or
    I.call("seaice_evp_b200_create", domain)                                                            # its host emulation)
    I.call("seaice_evp_b200_update", domain)
    I.call("seaice_evp_b200_subcycle", domain)
    assert [n for n, _ in bridge.returns] == ["evp_create", "evp_set_options", "evp_update_step", "evp_run_subcycles", "evp_fetch"]
    assert all(rc == 0 for _, rc in bridge.returns), bridge.returns
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    cm, vm = common.masks_for(mesh, step)
    for k in common.COMPARE_CELL:
        assert np.array_equal(work[k][cm], ref[k][cm]), k
    for k in common.COMPARE_VERTEX:
        assert np.array_equal(work[k][vm], ref[k][vm]), k
    assert np.abs(ref["uVelocity"][vm]).max() > 0
    I.call("seaice_evp_b200_destroy", 0)
    assert bridge.returns[-1] == ("evp_destroy", 0) and I.globals["evphandle"] is None
