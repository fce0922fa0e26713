"""Why is bench.py's e2e slower than scripts/e2e_timing.py? Variants of the same loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SPERR_B200_TIMING"] = "1"
import numpy as np, torch
import sperr_b200, bench
variant = sys.argv[1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if "nccl" in variant:
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = "29511"
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    dist.barrier()
L = sperr_b200.load()
n = 1024
dims = (n, n, n)
vol = bench.field_torch(dims, (0, 0, 0), dev)
if "devfirst" in variant:
    from sperr_b200 import sharded
    for _ in range(2):
        s = sharded.compress_3d_sharded(L.lib, vol.view(n, n, n), dims, (256,) * 3, 3, 1e-3)
        b, sh = sharded.decompress_3d_sharded(L.lib, s, dev, True)
hvol = vol.cpu().pin_memory().numpy()
if "freevol" in variant:
    del vol; torch.cuda.empty_cache()
print("variant", variant, flush=True)
out = None
for it in range(3):
    t0 = time.perf_counter()
    rc, s = L.compress_3d(hvol, dims, (256,) * 3, 3, 1e-3, copy=False)
    t1 = time.perf_counter()
    if "keep" in variant:
        rc, out2, d = L.decompress_3d(s, True, copy=False)
        out = out2
    else:
        out = None
        rc, out, d = L.decompress_3d(s, True, copy=False)
    t2 = time.perf_counter()
    print("compress %.1f ms  decompress %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
