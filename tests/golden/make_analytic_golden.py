"""Golden vectors made by the REFERENCE'S OWN Python test-case scripts (imported from /root/reference, which
exists only in the build container -- this script is committed, its output is what travels):

  analytic_planar_hex82.npz   testing_and_setup/testcases/square/operators_strain_stress_divergence/create_ics.py
                              velocities_strains_stress_divergences(x, y)  (:12-48), A = B = C = D = 2.56
  analytic_sphere_ico5.npz    testing_and_setup/testcases/spherical_operators/strain_stress_divergence/create_ic.py
                              velocities_strains_analytical(lat, lon, 3, 5, 2, 4)  (:574-603), unit sphere,
                              rotated grid (grid_rotation_forward :607-623, latlon_from_xyz :627-634)

evaluated at the vertices of the meshes our generators produce for BASELINE configs[0] (planar hex 82 x 94,
dc = 0.0125) and for the reference's 10 242-cell sphere.  These are the analytic known answers the reference's
operator tests compare the model against (strain_stress_divergence_scaling.py:9-27); tests/test_analytic_golden.py
holds the oracle (CPU) and the device (GPU) to them.

    python tests/golden/make_analytic_golden.py        # needs /root/reference
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/testing_and_setup/testcases"


def _load(path, name):
    """Import a reference script; netCDF4 (absent here) is only used by its file-writing driver functions."""
    if "netCDF4" not in sys.modules:
        stub = types.ModuleType("netCDF4")
        stub.Dataset = None
        sys.modules["netCDF4"] = stub
    import scipy.special
    if not hasattr(scipy.special, "sph_harm"):      # imported by the sphere script but never called (it brings its
        scipy.special.sph_harm = None               # own Legendre polynomials); removed from recent SciPy
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def planar():
    from mpas_seaice_b200 import meshgen
    ref = _load(os.path.join(REF, "square/operators_strain_stress_divergence/create_ics.py"), "ref_square_ics")
    mesh = meshgen.planar_hex(82, 94, 0.0125)
    nV = mesh.nVertices
    cols = [[] for _ in range(7)]
    for i in range(nV):
        vals = ref.velocities_strains_stress_divergences(float(mesh.xVertex[i]), float(mesh.yVertex[i]))
        for c, v in zip(cols, vals):
            c.append(v)
    out = dict(zip(("u", "v", "e11", "e22", "e12", "divu", "divv"), (np.array(c) for c in cols)))
    out.update(x=mesh.xVertex[:nV].copy(), y=mesh.yVertex[:nV].copy(), nx=np.int64(82), ny=np.int64(94), dc=np.float64(0.0125))
    np.savez_compressed(os.path.join(HERE, "analytic_planar_hex82.npz"), **out)
    print("planar", nV, "vertices")


def sphere():
    from mpas_seaice_b200 import meshgen
    ref = _load(os.path.join(REF, "spherical_operators/strain_stress_divergence/create_ic.py"), "ref_sphere_ic")
    import math
    # the script calls math.factorial with integral floats (fabs(m)), which Python >= 3.12 rejects: same values
    ref.factorial = lambda x: math.factorial(int(x))
    mesh = meshgen.icosphere(5, radius=1.0)
    nV = mesh.nVertices
    cols = [[] for _ in range(7)]
    lats, lons = [], []
    for i in range(nV):
        xp, yp, zp = ref.grid_rotation_forward(float(mesh.xVertex[i]), float(mesh.yVertex[i]), float(mesh.zVertex[i]), True)
        lat, lon = ref.latlon_from_xyz(xp, yp, zp, 1.0)
        vals = ref.velocities_strains_analytical(lat, lon, 3, 5, 2, 4)
        lats.append(lat)
        lons.append(lon)
        for c, v in zip(cols, vals):
            c.append(np.real(v))
    out = dict(zip(("u", "v", "e11", "e22", "e12", "divu", "divv"), (np.array(c, dtype=np.float64) for c in cols)))
    out.update(x=mesh.xVertex[:nV].copy(), y=mesh.yVertex[:nV].copy(), z=mesh.zVertex[:nV].copy(),
               latRotated=np.array(lats), lonRotated=np.array(lons), level=np.int64(5))
    np.savez_compressed(os.path.join(HERE, "analytic_sphere_ico5.npz"), **out)
    print("sphere", nV, "vertices")


if __name__ == "__main__":
    planar()
    sphere()
