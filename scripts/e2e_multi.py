"""Plain sperr_comp_3d / sperr_decomp_3d (host buffers, pageable input) on the bench volume with one
GPU and with SPERR_B200_DEVICES=all: the reference-facing call spreading one volume over the GPUs of
the box (csrc/capi_multi.cu). Prints times and checks that the containers and the decoded bits agree."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dims = (n, n, n)
vol = bench.field_torch(dims, (0, 0, 0), torch.device("cuda", 0)).cpu().numpy().copy()   # pageable
ref_s = ref_o = None
for devs in (None, "all"):
    if devs:
        os.environ["SPERR_B200_DEVICES"] = devs
    else:
        os.environ.pop("SPERR_B200_DEVICES", None)
    best = [1e9, 1e9]
    for it in range(4):
        t0 = time.perf_counter()
        rc, s = L.compress_3d(vol, dims, (256,) * 3, 3, 1e-3, copy=False)
        t1 = time.perf_counter()
        assert rc == 0, rc
        rc, out, d = L.decompress_3d(s, True, copy=False)
        t2 = time.perf_counter()
        assert rc == 0, rc
        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
        if ref_s is None:
            ref_s, ref_o = np.array(s, copy=True), np.array(out, copy=True)
        elif it == 0:
            assert np.array_equal(ref_s, s), "container differs from the one-GPU container"
            assert np.array_equal(ref_o.view(np.uint32), np.asarray(out).view(np.uint32)), "decoded bits differ"
        del out, s
    gb = vol.nbytes / 1e9
    print("devices=%s (%d visible): compress %.1f ms  decompress %.1f ms  e2e %.2f GB/s" % (
        devs or "one", torch.cuda.device_count(), best[0] * 1e3, best[1] * 1e3, gb / (best[0] + best[1])), flush=True)
