"""BASELINE.json config #5 (batched 2D: slices of 2048^2 fp32, PWE) on one GPU, device-resident:
compress / decompress input GB/s through the batched slice entry points, and the reference's
sperr_comp_2d / sperr_decomp_2d on a few slices (one slice per call, as its API works).
    python scripts/bench2d.py [nslices] [dim]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda", 0)
vol = bench.field_torch((dim, dim, ns), (0, 0, 0), dev)   # ns z-planes = ns slices
nbytes = vol.numel() * 4
tol = 1e-3
res = {"workload": "%d slices of %dx%d fp32, PWE %g" % (ns, dim, dim, tol)}
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc, streams, lens = L.compress_2d_batch(vol.data_ptr(), True, (dim, dim), ns, 3, tol, device=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    assert rc == 0
    out = torch.empty_like(vol)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    rc, _ = L.decompress_2d_batch(streams, lens, (dim, dim), True, d_out_ptr=out.data_ptr())
    torch.cuda.synchronize(); t3 = time.perf_counter()
    assert rc == 0
    res["compress_gbs"] = nbytes / (t1 - t0) / 1e9
    res["decompress_gbs"] = nbytes / (t3 - t2) / 1e9
    res["compress_ms"] = (t1 - t0) * 1e3
    res["decompress_ms"] = (t3 - t2) * 1e3
res["max_abs_err"] = float((out.double() - vol.double()).abs().max())
res["bpp"] = float(lens.sum()) * 8 / vol.numel()
# reference: one slice per call
import refs
R = refs.ref()
if R is not None:
    k = min(ns, 4)
    h = vol.view(ns, dim, dim)[:k].cpu().numpy()
    t0 = time.perf_counter()
    ss = [R.comp_2d(h[i], (dim, dim), 3, tol)[1] for i in range(k)]
    t1 = time.perf_counter()
    dd = [R.decomp_2d(s, (dim, dim), True)[1] for s in ss]
    t2 = time.perf_counter()
    res["ref_compress_gbs"] = k * dim * dim * 4 / (t1 - t0) / 1e9
    res["ref_decompress_gbs"] = k * dim * dim * 4 / (t2 - t1) / 1e9
    off = 0
    same = True
    for i in range(k):
        l = int(lens[i])
        same &= bool(np.array_equal(np.asarray(streams[off:off + l]), ss[i]))
        off += l
    res["streams_equal_reference_first_%d" % k] = same
print(json.dumps(res))
