"""Prints the fast decoder's per-phase cycle totals for one 256^3 chunk of the bench workload, and
times the host-buffer C API calls (compress / decompress) separately."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SPERR_B200_DECPROF"] = "1"
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dims = (n, n, n)
vol = bench.field_torch(dims, (0, 0, 0), torch.device("cuda", 0)).cpu().numpy()
for it in range(2):
    t0 = time.perf_counter()
    rc, s = L.compress_3d(vol, dims, (256,) * 3, 3, 1e-3, copy=False)
    t1 = time.perf_counter()
    rc, out, d = L.decompress_3d(s, True, copy=False)
    t2 = time.perf_counter()
    print("n=%d compress %.1f ms  decompress %.1f ms  stream %d B" % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, s.size), flush=True)
