import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SPERR_B200_DECTRACE"] = "1"
os.environ.setdefault("SPERR_B200_DEC_CLUSTER", "1")
import numpy as np
import gpulib, refs
a = sys.argv[1] if len(sys.argv) > 1 else "cuda"
lib = gpulib.load(a) if a in ("cuda", "emul") else gpulib.Lib(a)
oracle = refs.oracle()
dims = (64, 64, 64)
v = refs.synthetic_field(dims, seed=5)
rc, s = oracle.comp_3d(v, dims, dims, 1, 3.0)
rc, got, d = lib.decomp_3d(s, True)
print("rc", rc)
