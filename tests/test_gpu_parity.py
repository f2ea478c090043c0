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
