"""debug: decompress a few power-of-two cases through the C API and compare with the oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpulib, refs, cases
so = sys.argv[1] if len(sys.argv) > 1 else gpulib.CUDA_SO
lib = gpulib.load(so) if so in ("cuda", "emul") else gpulib.Lib(so)
oracle = refs.oracle()
todo = [((64, 32, 16), (64, 32, 16), 3, 1e-4), ((64, 64, 64), (64, 64, 64), 1, 3.0), ((32, 32, 32), (16, 16, 16), 3, 1e-3),
        ((128, 128, 128), (128, 128, 128), 3, 1e-3)]
if os.environ.get('DEC_DEBUG_ONE'):
    todo = todo[:1]
for dims, chunks, mode, q in todo:
    v = refs.synthetic_field(dims, seed=5)
    rc, s = oracle.comp_3d(v, dims, chunks, mode, q)
    rc, got, d = lib.decomp_3d(s, True)
    rc2, exp, d2 = oracle.decomp_3d(s, True)
    bad = -1 if rc != 0 else int(np.count_nonzero(got.view(np.uint32) != exp.view(np.uint32)))
    print("case", dims, chunks, mode, q, "rc", rc, "differing values", bad, flush=True)
