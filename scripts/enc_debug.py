"""debug: compress small cases through the C API and compare with the oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpulib, refs
lib = gpulib.load(sys.argv[1] if len(sys.argv) > 1 else "cuda")
oracle = refs.oracle()
v17 = refs.load_test_data("wmag17.float")
todo = [(v17, (17, 17, 17), (17, 17, 17), 3, 0.3)]
for dims, chunks, mode, q in (((32, 32, 32), (16, 16, 16), 3, 1e-3), ((64, 64, 64), (64, 64, 64), 1, 3.0), ((128, 64, 64), (64, 64, 64), 3, 1e-3)):
    todo.append((refs.synthetic_field(dims, seed=5), dims, chunks, mode, q))
if os.environ.get("ENC_DEBUG_ONE"):
    todo = todo[:int(os.environ["ENC_DEBUG_ONE"])]
for v, dims, chunks, mode, q in todo:
    rc, got = lib.comp_3d(v, dims, chunks, mode, q)
    rc2, exp = oracle.comp_3d(v, dims, chunks, mode, q)
    same = rc == 0 and got.size == exp.size and bool(np.array_equal(got, exp))
    print("case", dims, chunks, mode, q, "rc", rc, "equal", same, flush=True)
