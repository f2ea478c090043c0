import os, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in ("", "tests", os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
from mpas_seaice_b200 import host
import evp_emu
host._lib = host.load_library(evp_emu.library())
import test_gpu_fuzz as F
import inspect
src = inspect.getsource(F.test_random_configurations_match_oracle)
bad = []
t0 = time.time()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    try:
        F.test_random_configurations_match_oracle.__wrapped__(host._lib, seed) if hasattr(F.test_random_configurations_match_oracle, "__wrapped__") else F.test_random_configurations_match_oracle(host._lib, seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
    except Exception as e:
        bad.append((seed, "EXC " + repr(e)[:200]))
print("seeds", lo, hi, "failures:", bad, "%.0fs" % (time.time() - t0))
