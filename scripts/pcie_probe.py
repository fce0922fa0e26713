"""Host<->device copy rates of this box: pinned and pageable, 1 GiB pieces."""
import time, torch, numpy as np
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
hq = torch.empty(n, dtype=torch.uint8); hq.fill_(1)
def t(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
print("H2D pinned   %.1f GB/s" % t(lambda: d.copy_(hp, non_blocking=True)))
print("D2H pinned   %.1f GB/s" % t(lambda: hp.copy_(d, non_blocking=True)))
print("H2D pageable %.1f GB/s" % t(lambda: d.copy_(hq)))
print("D2H pageable %.1f GB/s" % t(lambda: hq.copy_(d)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda"); hp2 = torch.empty(n, dtype=torch.uint8).pin_memory()
def both():
    with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2): hp2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrent, each %.1f GB/s" % t(both))
import os
print("cpus", os.cpu_count())
a = np.empty(n, dtype=np.uint8); b = np.ones(n, dtype=np.uint8)
t0 = time.perf_counter(); a[:] = b; t1 = time.perf_counter()
print("host memcpy 1 thread (first touch) %.1f GB/s" % (n / (t1 - t0) / 1e9))
t0 = time.perf_counter(); a[:] = b; t1 = time.perf_counter()
print("host memcpy 1 thread (warm) %.1f GB/s" % (n / (t1 - t0) / 1e9))
