"""Is pinning the caller's result buffer (cudaHostRegister) cheap enough to DMA straight into it?"""
import ctypes as C, time, torch
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
libc.free.argtypes = [C.c_void_p]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
libc.memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]; libc.memset.restype = C.c_void_p
torch.cuda.init()
_rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
class RT:
    def cudaHostRegister(self, p, n, f):
        _rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]; return _rt.cudaHostRegister(p, n, f)
    def cudaHostUnregister(self, p):
        _rt.cudaHostUnregister.argtypes = [C.c_void_p]; return _rt.cudaHostUnregister(p)
    def cudaMemcpy(self, d, s, n, k):
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]; return _rt.cudaMemcpy(d, s, n, k)
rt = RT()
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for thp in (1, 0, 1):
    p = libc.malloc(n + 4096)
    a = (p + 4095) & ~4095
    if thp:
        libc.madvise(a, n, 14)  # MADV_HUGEPAGE
    t0 = time.perf_counter(); libc.memset(a, 0, n); t1 = time.perf_counter()
    r = rt.cudaHostRegister(a, n, 0); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    rc = rt.cudaMemcpy(a, d.data_ptr(), n, 2)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    u = rt.cudaHostUnregister(a); t5 = time.perf_counter()
    libc.free(p); t6 = time.perf_counter()
    print("thp=%d touch %.0f ms  register %.1f ms (%s)  memcpy %.1f ms (%s)  unregister %.1f ms  free %.1f ms" % (
        thp, (t1 - t0) * 1e3, (t2 - t1) * 1e3, r, (t4 - t3) * 1e3, rc, (t5 - t4) * 1e3, (t6 - t5) * 1e3), flush=True)
# registering in 256 MiB pieces (could be pipelined with the DMA)
p = libc.malloc(n + 4096); a = (p + 4095) & ~4095
libc.madvise(a, n, 14); libc.memset(a, 0, n)
piece = 256 << 20
t0 = time.perf_counter()
for off in range(0, n, piece):
    rt.cudaHostRegister(a + off, piece, 0)
t1 = time.perf_counter()
for off in range(0, n, piece):
    rt.cudaHostUnregister(a + off)
t2 = time.perf_counter()
print("16 x 256 MiB: register %.1f ms  unregister %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
# untouched memory: register does the faulting
p2 = libc.malloc(n + 4096); a2 = (p2 + 4095) & ~4095
libc.madvise(a2, n, 14)
t0 = time.perf_counter(); r = rt.cudaHostRegister(a2, n, 0); t1 = time.perf_counter()
print("untouched thp: register %.1f ms (%s)" % ((t1 - t0) * 1e3, r))
