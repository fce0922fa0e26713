import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpulib, refs, cases
lib = gpulib.load("cuda"); oracle = refs.oracle()
case = ((100, 88, 40), (50, 44, 40), 3, 2e-3)
dims, chunks, mode, q = case
v = refs.synthetic_field(dims, seed=5)
rc, got = lib.comp_3d(v, dims, chunks, mode, q)
rc2, exp = oracle.comp_3d(v, dims, chunks, mode, q)
print("comp rc", rc, "equal", rc == 0 and np.array_equal(got, exp), flush=True)
rc, dec, d = lib.decomp_3d(exp, True)
rc2, dexp, d2 = oracle.decomp_3d(exp, True)
print("decomp rc", rc, "equal", rc == 0 and np.array_equal(dec.view(np.uint32), dexp.view(np.uint32)), flush=True)
