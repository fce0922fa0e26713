"""Phase timing of the host-pointer C API on the bench workload (pinned input, as bench.py's e2e)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SPERR_B200_TIMING"] = "1"
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dims = (n, n, n)
vol = bench.field_torch(dims, (0, 0, 0), torch.device("cuda", 0)).cpu().pin_memory().numpy()
for it in range(3):
    t0 = time.perf_counter()
    rc, s = L.compress_3d(vol, dims, (256,) * 3, 3, 1e-3, copy=False)
    t1 = time.perf_counter()
    rc, out, d = L.decompress_3d(s, True, copy=False)
    t2 = time.perf_counter()
    del out
    t3 = time.perf_counter()
    print("n=%d compress %.1f ms  decompress %.1f ms  free %.1f ms" % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
