"""Regenerates tests/golden/*.npz: small, fully self-contained input/output vectors of the hot path.

The reference ships no numeric outputs and cannot be built here, so these vectors are
produced by the ORACLE (oracle/evp_oracle.c, the plain-C restatement of the cited reference routines)
and pin it against silent drift: tests/test_golden.py replays them through the oracle (CPU) and through
libevp_b200.so (GPU) and demands the same bits.  Inputs are stored too, and the subcycle uses only
+ - * / sqrt, so the files are independent of the machine's libm.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import common  # noqa: E402

MESH_KEYS = ("nEdgesOnCell", "verticesOnCell", "cellsOnVertex", "areaCell")
VAR_KEYS = ("cellVerticesAtVertex", "basisGradientU", "basisGradientV", "basisIntegralsU", "basisIntegralsV",
            "basisIntegralsMetric", "tanLatVertexRotatedOverRadius", "variationalDenominator")
CASES = {
    # name: (mesh kind, constitutive relation, subcycles, extra options)
    "hex20_evp_120": ("hex20", "evp", 120, {}),
    "ico3_evp_120": ("ico3", "evp", 120, {}),
    "ico3_revised_40": ("ico3", "evp_revised", 40, {}),
    "quad40_evp_avg_30": ("quad40", "evp", 30, {"average_variational_strain": True}),
}


def build(name):
    kind, cr, nsub, extra = CASES[name]
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh, constitutive_relation_type=cr)
    opts = dict(opts, **extra)
    ref = common.run_oracle(mesh, var, step, opts, nsub)
    out = {"nsub": np.int64(nsub)}
    for k in ("nCells", "nVertices", "maxEdges", "vertexDegree"):
        out["mesh_" + k] = np.int64(mesh[k])
    for k in MESH_KEYS:
        out["mesh_" + k] = mesh[k]
    for k in VAR_KEYS:
        out["var_" + k] = var[k]
    for k, v in step.items():
        if isinstance(v, np.ndarray):
            out["in_" + k] = v
    for k, v in opts.items():
        out["opt_" + k] = np.array(v)
    for k in common.COMPARE_CELL + common.COMPARE_VERTEX:
        out["out_" + k] = ref[k]
    return out


if __name__ == "__main__":
    for name in CASES:
        data = build(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path) // 1024, "KiB")
