"""How far one dynamics step (120 EVP subcycles) moves when the arithmetic is perturbed at the last bit -- the
numbers quoted in DESIGN.md ("FP64 reproducibility").  CPU only: the oracle built twice.

  1. FMA contraction: oracle built with -ffp-contract=fast -mfma against the shipped -ffp-contract=off build;
  2. basisIntegralsMetric stored as a symmetric matrix (its computed asymmetry is 2.6e-16 relative).

north_star allows 1e-10 relative max-norm on velocities and stresses after one dynamics step.
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import oracle  # noqa: E402


def report(tag, out, ref, cm, vm):
    vals = {k: common.rel_max_err(out[k], ref[k], vm) for k in ("uVelocity", "vVelocity")}
    vals.update({k: common.rel_max_err(out[k], ref[k], cm) for k in ("stress11", "stress22", "stress12")})
    print(tag, " ".join(f"{k}={v:.2e}" for k, v in vals.items()))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "ico5"
    mesh, var = common.mesh_case(kind)
    step, opts = common.step_case(mesh)
    cm, vm = common.masks_for(mesh, step)
    ref = common.run_oracle(mesh, var, step, opts, 120)

    sym = dict(var)
    sM = var["basisIntegralsMetric"].copy()
    iu = np.triu_indices(sM.shape[1], 1)
    sM[:, iu[1], iu[0]] = sM[:, iu[0], iu[1]]
    sym["basisIntegralsMetric"] = sM
    report("symmetric metric integrals:", common.run_oracle(mesh, sym, step, opts, 120), ref, cm, vm)

    with tempfile.TemporaryDirectory() as tmp:
        lib = os.path.join(tmp, "liboracle_fma.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("evp_oracle.c", "evp_precompute_oracle.c", "ir_oracle.c")]
        subprocess.run(["gcc", "-O2", "-ffp-contract=fast", "-mfma", "-fno-fast-math", "-fPIC", "-fopenmp", "-shared", "-o", lib]
                       + srcs + ["-lm"], check=True)
        oracle._LIB_PATH, oracle._lib = lib, None
        report("FMA contraction:           ", common.run_oracle(mesh, var, step, opts, 120), ref, cm, vm)
    assert isinstance(oracle.lib(), ctypes.CDLL)


if __name__ == "__main__":
    main()
