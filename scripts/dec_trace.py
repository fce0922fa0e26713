import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SPERR_B200_DECTRACE"] = "1"
os.environ.setdefault("SPERR_B200_DEC_CLUSTER", "1")
import numpy as np
import gpulib, refs, bench
a = sys.argv[1] if len(sys.argv) > 1 else "cuda"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = gpulib.load(a) if a in ("cuda", "emul") else gpulib.Lib(a)
oracle = refs.oracle()
if n:
    dims = (n, n, n)
    v = bench.field_numpy(dims)
    rc, s = oracle.comp_3d(v, dims, dims, 3, 1e-3)
else:
    dims = (64, 64, 64)
    v = refs.synthetic_field(dims, seed=5)
    rc, s = oracle.comp_3d(v, dims, dims, 1, 3.0)
rc, got, d = lib.decomp_3d(s, True)
rc2, exp, d2 = oracle.decomp_3d(s, True)
bad = -1 if rc != 0 else int(np.count_nonzero(got.view(np.uint32) != exp.view(np.uint32)))
print("rc", rc, "differing", bad)
