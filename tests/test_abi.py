"""The C-ABI library: loads, exports every symbol include/evp_b200.h declares, mirrors the header's struct
layouts, and FAILS LOUDLY (no CPU fallback) when no CUDA device answers.  No compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mpas_seaice_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(evp_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(evp_lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(evp_lib, n), f"{n} declared in include/evp_b200.h but not exported"
    assert set(host.EXPORTS) == set(names), set(host.EXPORTS) ^ set(names)


def test_struct_mirrors_match_header():
    src = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(name):
        body = re.search(r"typedef struct \{([^{}]*)\} " + name + ";", src).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                out.append(re.findall(r"[A-Za-z_0-9]+", decl)[-1])
        return out

    assert fields("evp_mesh_desc") == [f[0] for f in host.MeshDesc._fields_]
    assert fields("evp_options") == [f[0] for f in host.Options._fields_]
    assert fields("evp_step_fields") == list(host.STEP_FIELDS)
    assert fields("evp_out_fields") == list(host.OUT_FIELDS)


def test_argument_errors_do_not_need_a_device(evp_lib):
    h = C.c_void_p()
    md, o = host.MeshDesc(), host.make_options(dict(elasticTimeStep=30.0, dynamicsTimeStep=3600.0,
                                                    dampingTimescale=1296.0))
    assert evp_lib.evp_create(C.byref(h), None, C.byref(o)) == 1            # EVP_ERR_ARGUMENT
    assert b"NULL" in evp_lib.evp_last_error_string()
    o.constitutive_relation_type = 9
    assert evp_lib.evp_create(C.byref(h), C.byref(md), C.byref(o)) == 1
    assert evp_lib.evp_run_subcycles(None, 1) == 1
    assert evp_lib.evp_destroy(None) == 0
    # the widened entry points validate their arguments before touching the device too
    for fn in ("evp_set_mesh_ext", "evp_pre_subcycle", "evp_post_subcycle", "evp_fetch_pre", "evp_set_weak_mesh",
               "evp_update_weak_state", "evp_fetch_weak", "evp_release_host_memory"):
        f = getattr(evp_lib, fn)
        n_args = {"evp_pre_subcycle": 3, "evp_release_host_memory": 1}.get(fn, 2)
        assert f(*([None] * n_args)) == 1, fn
        assert b"NULL" in evp_lib.evp_last_error_string(), fn
    # scheme combinations (velocity_solver.F:195-198)
    o = host.make_options(dict(elasticTimeStep=30.0, dynamicsTimeStep=3600.0, dampingTimescale=1296.0,
                               strain_scheme="variational", stress_divergence_scheme="weak"))
    md.maxEdges, md.vertexDegree = 6, 3
    assert evp_lib.evp_create(C.byref(h), C.byref(md), C.byref(o)) == 1
    assert b"not a valid combination" in evp_lib.evp_last_error_string()


def test_host_metric_terms_is_scalar_libm():
    """evp_host_metric_terms (no device needed): tan(asin(z/R))/R element by element, independent of the position
    of the element in the array (the property numpy's SIMD loops do not have)."""
    import math
    rng = np.random.default_rng(5)
    R = 6371229.0
    z = R * np.sin(rng.uniform(-1.5, 1.5, 1001))
    out = host.host_metric_terms(z, R)
    want = np.array([math.tan(math.asin(v / R)) / R for v in z])
    assert np.array_equal(out, want)
    assert np.array_equal(host.host_metric_terms(z[3:40].copy(), R), out[3:40])


def test_no_cpu_fallback_without_a_device(evp_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from mpas_seaice_b200 import meshgen, variational_init
    mesh = meshgen.planar_hex(6, 6, 1000.0)
    static = variational_init.init_static(mesh)
    opts = dict(elasticTimeStep=30.0, dynamicsTimeStep=3600.0, dampingTimescale=1296.0)
    with pytest.raises(host.EvpError, match="error 2"):                     # EVP_ERR_CUDA
        host.EvpSolver(mesh, static, opts, local_coords=(static["xLocal"], static["yLocal"]))


def test_missing_library_raises(tmp_path):
    with pytest.raises(host.EvpError, match="no CPU fallback"):
        host.load_library(str(tmp_path / "libevp_b200.so"))
