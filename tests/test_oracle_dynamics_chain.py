"""The two oracles chained the way seaice_run_dynamics chains the two paths
(src/shared/mpas_seaice_time_integration.F:147-165: velocity solver, then advection with the new velocities):
a dynamics step of the EVP oracle on the sphere, its vertex velocities handed to the transport oracle.  Pins the
interface between the paths -- the same (nVertices+1) east / north velocity arrays, the same mesh conventions -- and
that the transported state stays physical under a solver-produced (not synthetic) velocity field."""
import numpy as np

import common
import oracle
from oracle import ir
from mpas_seaice_b200 import irmesh, synthetic


def test_evp_velocities_drive_the_transport():
    mesh, var = common.mesh_case("ico4")
    nC, nV = mesh.nCells, mesh.nVertices
    state = synthetic.sphere_state(mesh, "B")                     # ice caps, open ocean in between
    irf = irmesh.ir_fields(mesh)
    geom = ir.init_geometry(mesh, irf)
    tracers = ir.default_tracers(nC, 1)
    tracers[0].array[:nC, 0, 0] = state["iceAreaCell"][:nC] * 0.9
    tracers[1].array[:nC, 0, 0] = state["iceVolumeCell"][:nC] * 0.9
    tracers[3].array[:nC, 0, 0] = -10.0
    A = mesh.areaCell[:nC]
    area0, vol0 = (tracers[0].array[:nC, 0, 0] * A).sum(), (tracers[1].array[:nC, 0, 0] * A).sum()
    dt = 3600.0
    moved = 0.0
    for _ in range(3):
        # velocity solve from the CURRENT ice state (aggregate of the one category), zero initial stress
        st = dict(state)
        st["iceAreaCell"] = np.append(tracers[0].array[:nC, 0, 0], 0.0)
        st["iceVolumeCell"] = np.append(tracers[1].array[:nC, 0, 0], 0.0)
        step = oracle.pre_subcycle(mesh, st, dt)
        _, opts = synthetic.pre_subcycle(mesh, st, dt)
        oracle.subcycle_velocity_solver(mesh, var, step, opts, 120)
        u, v = step["uVelocity"], step["vVelocity"]
        assert u.shape == (nV + 1,) and np.isfinite(u).all() and np.isfinite(v).all()
        speed = np.hypot(u[:nV], v[:nV]).max()
        assert 0.0 < speed * dt < geom["minLengthEdgesOnVertex"][:nV].min()      # the transport's CFL condition
        before = tracers[0].array.copy()
        d = ir.run(mesh, irf, geom, tracers, u, v, dt, diagnostics=True)
        assert d["error"] == 0
        moved += np.abs(tracers[0].array - before).max()
    a, vol = tracers[0].array[:nC, 0, 0], tracers[1].array[:nC, 0, 0]
    assert moved > 1e-4                                           # the ice edge moved
    assert abs((a * A).sum() / area0 - 1) < 1e-13 and abs((vol * A).sum() / vol0 - 1) < 1e-13
    assert a.min() >= 0.0 and vol.min() >= 0.0
    ice = a > 1e-11
    assert np.allclose(tracers[3].array[:nC, 0, 0][ice], -10.0, rtol=1e-12)
    h = vol[ice] / a[ice]
    assert h.min() > 0.99 * 1.0 - 1e-9 and h.max() < 1.0 + 1e-9   # thickness 1 m everywhere it started: stays 1 m
