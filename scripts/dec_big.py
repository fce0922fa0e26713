"""debug: one bench-shaped chunk (or several) decoded with the given cluster size, against the oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpulib, refs, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nz = int(sys.argv[2]) if len(sys.argv) > 2 else n
lib = gpulib.load("cuda")
oracle = refs.oracle()
dims = (n, n, nz)
ck = (min(n, 256),) * 3
v = bench.field_numpy(dims)
rc, s = oracle.comp_3d(v, dims, ck, 3, 1e-3)
rc, got, d = lib.decomp_3d(s, True)
rc2, exp, d2 = oracle.decomp_3d(s, True)
bad = -1 if rc != 0 else int(np.count_nonzero(got.view(np.uint32) != exp.view(np.uint32)))
print("dims", dims, "R", os.environ.get("SPERR_B200_DEC_CLUSTER"), "rc", rc, "differing values", bad, flush=True)
