"""ctypes face of the product library's C ABI (include/sperr_b200.h).

``load("cuda")`` loads sperr_b200/libsperr_b200.so (the nvcc build; needs a GPU to run anything).
``load("emul")`` loads tests/emul/libsperr_emul.so, the same sources compiled against the CPU SIMT
emulator -- a debugging aid for the build container, never a product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_SO = os.path.join(ROOT, "sperr_b200", "libsperr_b200.so")
EMUL_SO = os.path.join(ROOT, "tests", "emul", "libsperr_emul.so")

sz = C.c_size_t
vp = C.c_void_p

_libc = C.CDLL(None)
_libc.free.argtypes = [vp]
_libc.free.restype = None


def _ptr(a):
    return a.ctypes.data_as(vp)


class Lib:
    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)

    def set_fma_flavour(self, on):
        f = self.lib.sperr_b200_set_fma_flavour
        f.restype = None
        f.argtypes = [C.c_int]
        f(int(on))

    # ---- stage hooks ----
    def stage_condition(self, vol, dims):
        vol = np.ascontiguousarray(vol)
        n = int(np.prod(dims))
        out = np.zeros(n, dtype=np.float64)
        mean = C.c_double(0)
        isc = C.c_int(0)
        f = self.lib.sperr_b200_stage_condition
        f.restype = C.c_int
        f.argtypes = [vp, C.c_int, sz, sz, sz, vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        rc = f(_ptr(vol), int(vol.dtype == np.float32), *dims, _ptr(out), C.byref(mean), C.byref(isc))
        assert rc == 0
        return out, mean.value, bool(isc.value)

    def stage_dwt(self, vals, dims, inverse=False, is_2d=False):
        buf = np.ascontiguousarray(vals, dtype=np.float64).copy()
        f = self.lib.sperr_b200_stage_dwt
        f.restype = C.c_int
        f.argtypes = [vp, sz, sz, sz, C.c_int, C.c_int]
        d = tuple(dims) + (1,) * (3 - len(dims))
        rc = f(_ptr(buf), *d, int(inverse), int(is_2d))
        assert rc == 0
        return buf

    def stage_dwt_fused(self, vals, dims, inverse=False):
        buf = np.ascontiguousarray(vals, dtype=np.float64).copy()
        f = self.lib.sperr_b200_stage_dwt_fused
        f.restype = C.c_int
        f.argtypes = [vp, sz, sz, sz, C.c_int]
        rc = f(_ptr(buf), *dims, int(inverse))
        return rc, buf

    def stage_quantize(self, vals, dims, q):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        mags = np.zeros(vals.size, dtype=np.uint64)
        signs = np.zeros(vals.size, dtype=np.uint8)
        wide = C.c_int(0)
        f = self.lib.sperr_b200_stage_quantize
        f.restype = C.c_int
        f.argtypes = [vp, sz, sz, sz, C.c_double, vp, vp, C.POINTER(C.c_int)]
        rc = f(_ptr(vals), *dims, q, _ptr(mags), _ptr(signs), C.byref(wide))
        return rc, mags, signs, wide.value

    def stage_speck3d_encode(self, mags, signs, dims, budget_bits=0):
        mags = np.ascontiguousarray(mags, dtype=np.uint64)
        signs = np.ascontiguousarray(signs, dtype=np.uint8)
        cap = mags.size * 10 + 64
        out = np.zeros(cap, dtype=np.uint8)
        n = sz(0)
        f = self.lib.sperr_b200_stage_speck3d_encode
        f.restype = C.c_int
        f.argtypes = [vp, vp, sz, sz, sz, sz, vp, sz, C.POINTER(sz)]
        rc = f(_ptr(mags), _ptr(signs), *dims, budget_bits, _ptr(out), cap, C.byref(n))
        assert rc == 0, rc
        return out[:n.value].copy()

    def stage_speck3d_decode(self, stream, dims):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        n = int(np.prod(dims))
        mags = np.zeros(n, dtype=np.uint64)
        signs = np.zeros(n, dtype=np.uint8)
        f = self.lib.sperr_b200_stage_speck3d_decode
        f.restype = C.c_int
        f.argtypes = [vp, sz, sz, sz, sz, vp, vp]
        rc = f(_ptr(stream), stream.size, *dims, _ptr(mags), _ptr(signs))
        assert rc == 0, rc
        return mags, signs

    def stage_outlier_encode(self, pos, err, total_len, tol):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        cap = total_len // 2 + pos.size * 16 + 64
        out = np.zeros(cap, dtype=np.uint8)
        n = sz(0)
        f = self.lib.sperr_b200_stage_outlier_encode
        f.restype = C.c_int
        f.argtypes = [vp, vp, sz, sz, C.c_double, vp, sz, C.POINTER(sz)]
        rc = f(_ptr(pos), _ptr(err), pos.size, total_len, tol, _ptr(out), cap, C.byref(n))
        assert rc == 0, rc
        return out[:n.value].copy()

    def stage_outlier_decode(self, stream, total_len, tol):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        cap = total_len
        pos = np.zeros(cap, dtype=np.uint64)
        err = np.zeros(cap, dtype=np.float64)
        n = sz(0)
        f = self.lib.sperr_b200_stage_outlier_decode
        f.restype = C.c_int
        f.argtypes = [vp, sz, sz, C.c_double, vp, vp, sz, C.POINTER(sz)]
        rc = f(_ptr(stream), stream.size, total_len, tol, _ptr(pos), _ptr(err), cap, C.byref(n))
        assert rc == 0, rc
        return pos[:n.value].copy(), err[:n.value].copy()

    # ---- reference-compatible API ----
    def comp_3d(self, vol, dims, chunks, mode, quality, nthreads=0):
        vol = np.ascontiguousarray(vol)
        assert vol.dtype in (np.float32, np.float64)
        f = self.lib.sperr_comp_3d
        f.restype = C.c_int
        f.argtypes = [vp, C.c_int] + [sz] * 6 + [C.c_int, C.c_double, sz, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None)
        n = sz(0)
        rc = f(_ptr(vol), int(vol.dtype == np.float32), *dims, *chunks, mode, quality, nthreads,
               C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        _libc.free(dst)
        return 0, out

    def decomp_3d(self, stream, output_float=True, nthreads=0):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self.lib.sperr_decomp_3d
        f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_int, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(vp)]
        dx, dy, dz = sz(0), sz(0), sz(0)
        dst = vp(None)
        rc = f(_ptr(stream), stream.size, int(output_float), nthreads, C.byref(dx), C.byref(dy),
               C.byref(dz), C.byref(dst))
        if rc != 0:
            return rc, None, None
        dt = np.float32 if output_float else np.float64
        n = dx.value * dy.value * dz.value
        out = np.frombuffer(C.string_at(dst.value, n * np.dtype(dt).itemsize), dtype=dt).copy()
        _libc.free(dst)
        return 0, out, (dx.value, dy.value, dz.value)


    def trunc_3d(self, stream, pct):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self.lib.sperr_trunc_3d
        f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_uint, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None)
        n = sz(0)
        rc = f(_ptr(stream), stream.size, pct, C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        _libc.free(dst)
        return 0, out

    def decomp_3d_multires(self, stream, output_float=False):
        """-> (rc, full volume, dims, [coarse volumes, coarsest first], [their dims])"""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self.lib.sperr_b200_decomp_3d_multires
        f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_int, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(vp),
                      C.POINTER(sz), vp, vp]
        dx, dy, dz, nl = sz(0), sz(0), sz(0), sz(0)
        dst = vp(None)
        ld = np.zeros(24, dtype=np.uint64)
        lv = (vp * 8)()
        rc = f(_ptr(stream), stream.size, int(output_float), C.byref(dx), C.byref(dy), C.byref(dz),
               C.byref(dst), C.byref(nl), _ptr(ld), C.cast(lv, vp))
        if rc != 0:
            return rc, None, None, None, None
        dt = np.float32 if output_float else np.float64
        isz = np.dtype(dt).itemsize
        n = dx.value * dy.value * dz.value
        full = np.frombuffer(C.string_at(dst.value, n * isz), dtype=dt).copy()
        _libc.free(dst)
        levels, dims = [], []
        for h in range(nl.value):
            d = tuple(int(x) for x in ld[3 * h:3 * h + 3])
            m = d[0] * d[1] * d[2]
            levels.append(np.frombuffer(C.string_at(lv[h], m * isz), dtype=dt).copy())
            dims.append(d)
            _libc.free(lv[h])
        return 0, full, (dx.value, dy.value, dz.value), levels, dims

    # ---- 2D slices ----
    def stage_speck2d_encode(self, mags, signs, dims, budget_bits=0):
        mags = np.ascontiguousarray(mags, dtype=np.uint64)
        signs = np.ascontiguousarray(signs, dtype=np.uint8)
        cap = mags.size * 10 + 64
        out = np.zeros(cap, dtype=np.uint8)
        n = sz(0)
        f = self.lib.sperr_b200_stage_speck2d_encode
        f.restype = C.c_int
        f.argtypes = [vp, vp, sz, sz, sz, vp, sz, C.POINTER(sz)]
        rc = f(_ptr(mags), _ptr(signs), dims[0], dims[1], budget_bits, _ptr(out), cap, C.byref(n))
        assert rc == 0, rc
        return out[:n.value].copy()

    def comp_2d(self, img, dims, mode, quality, header=False):
        img = np.ascontiguousarray(img)
        f = self.lib.sperr_comp_2d
        f.restype = C.c_int
        f.argtypes = [vp, C.c_int, sz, sz, C.c_int, C.c_double, C.c_int, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None)
        n = sz(0)
        rc = f(_ptr(img), int(img.dtype == np.float32), dims[0], dims[1], mode, quality, int(header),
               C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        _libc.free(dst)
        return 0, out

    def decomp_2d(self, stream, dims, output_float=True):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self.lib.sperr_decomp_2d
        f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_int, sz, sz, C.POINTER(vp)]
        dst = vp(None)
        rc = f(_ptr(stream), stream.size, int(output_float), dims[0], dims[1], C.byref(dst))
        if rc != 0:
            return rc, None
        dt = np.float32 if output_float else np.float64
        n = dims[0] * dims[1]
        out = np.frombuffer(C.string_at(dst.value, n * np.dtype(dt).itemsize), dtype=dt).copy()
        _libc.free(dst)
        return 0, out

    def comp_2d_batch(self, imgs, dims, mode, quality, header=False):
        """imgs: (nslices, dimy, dimx) array -> list of per-slice streams"""
        imgs = np.ascontiguousarray(imgs)
        ns = imgs.size // (dims[0] * dims[1])
        f = self.lib.sperr_b200_comp_2d_batch
        f.restype = C.c_int
        f.argtypes = [vp, C.c_int, sz, sz, sz, C.c_int, C.c_double, C.c_int, C.POINTER(vp), vp]
        dst = vp(None)
        lens = np.zeros(ns, dtype=np.uint64)
        rc = f(_ptr(imgs), int(imgs.dtype == np.float32), dims[0], dims[1], ns, mode, quality,
               int(header), C.byref(dst), _ptr(lens))
        if rc != 0:
            return rc, None
        total = int(lens.sum())
        buf = np.frombuffer(C.string_at(dst.value, total), dtype=np.uint8).copy()
        _libc.free(dst)
        out, off = [], 0
        for l in lens:
            out.append(buf[off:off + int(l)])
            off += int(l)
        return 0, out

    def decomp_2d_batch(self, streams, dims, output_float=True):
        lens = np.array([s.size for s in streams], dtype=np.uint64)
        buf = np.ascontiguousarray(np.concatenate(streams), dtype=np.uint8)
        f = self.lib.sperr_b200_decomp_2d_batch
        f.restype = C.c_int
        f.argtypes = [vp, vp, sz, C.c_int, sz, sz, C.POINTER(vp)]
        dst = vp(None)
        rc = f(_ptr(buf), _ptr(lens), len(streams), int(output_float), dims[0], dims[1], C.byref(dst))
        if rc != 0:
            return rc, None
        dt = np.float32 if output_float else np.float64
        n = dims[0] * dims[1] * len(streams)
        out = np.frombuffer(C.string_at(dst.value, n * np.dtype(dt).itemsize), dtype=dt).copy()
        _libc.free(dst)
        return 0, out.reshape(len(streams), dims[1], dims[0])


def build_emul():
    subprocess.run([os.path.join(ROOT, "tests", "emul", "build.sh")], check=True,
                   stdout=subprocess.DEVNULL)


def load(kind):
    if kind == "emul":
        build_emul()
        return Lib(EMUL_SO)
    if not os.path.exists(CUDA_SO):
        raise RuntimeError("sperr_b200/libsperr_b200.so is missing: run __graft_entry__.build()")
    return Lib(CUDA_SO)
