"""ctypes bindings for the two CPU checkers used by the tests.

* ``oracle``  -- oracle/liboracle.so, our plain-C restatement (always buildable).
* ``ref``     -- oracle/_ref/libsperr_ref.so, the unmodified reference compiled by oracle/Makefile
                 (present when built in the container that has /root/reference; it travels to the
                 GPU box with the repo snapshot).

Test infrastructure only: nothing under sperr_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsperr_ref.so")

sz = C.c_size_t
vp = C.c_void_p
u8p = C.POINTER(C.c_uint8)


def _build(target):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target], check=False,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def load_oracle():
    if not os.path.exists(ORACLE_SO) or (os.path.getmtime(ORACLE_SO) <
                                         os.path.getmtime(os.path.join(ROOT, "oracle", "sperr_oracle.c"))):
        _build("oracle")
    return C.CDLL(ORACLE_SO)


def load_ref():
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        _build("ref")
    if not os.path.exists(REF_SO):
        return None
    return C.CDLL(REF_SO)


def _ptr(a):
    return a.ctypes.data_as(vp)


class Coder:
    """Uniform python face over either checker library (prefix 'so_' or 'ref_')."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        self.is_ref = prefix == "ref_"

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- geometry ----
    def num_of_xforms(self, n):
        f = self._f("num_of_xforms"); f.restype = sz; f.argtypes = [sz]
        return f(n)

    def num_of_partitions(self, n):
        f = self._f("num_of_partitions"); f.restype = sz; f.argtypes = [sz]
        return f(n)

    def can_use_dyadic(self, nx, ny, nz):
        f = self._f("can_use_dyadic"); f.restype = C.c_int; f.argtypes = [sz] * 3
        return f(nx, ny, nz)

    def calc_approx_detail_len(self, n, lev):
        f = self._f("calc_approx_detail_len"); f.restype = None; f.argtypes = [sz, sz, vp]
        out = np.zeros(2, dtype=np.uint64)
        f(n, lev, _ptr(out))
        return int(out[0]), int(out[1])

    def chunk_volume(self, vol, chunk):
        f = self._f("chunk_volume"); f.restype = sz; f.argtypes = [sz] * 6 + [vp, sz]
        n = f(*vol, *chunk, None, 0)
        out = np.zeros((n, 6), dtype=np.uint64)
        f(*vol, *chunk, _ptr(out), n)
        return out

    # ---- stages ----
    def condition(self, vals, dims):
        buf = np.ascontiguousarray(vals, dtype=np.float64).copy()
        hdr = np.zeros(17, dtype=np.uint8)
        if self.is_ref:
            f = self.lib.ref_condition; f.restype = None; f.argtypes = [vp, sz, sz, sz, vp]
            f(_ptr(buf), *dims, _ptr(hdr))
        else:
            f = self.lib.so_condition; f.restype = C.c_int; f.argtypes = [vp, sz, vp]
            f(_ptr(buf), buf.size, _ptr(hdr))
        return buf, hdr

    def dwt3d(self, vals, dims, inverse=False):
        buf = np.ascontiguousarray(vals, dtype=np.float64).copy()
        if self.is_ref:
            f = self.lib.ref_dwt3d; f.restype = None; f.argtypes = [vp, sz, sz, sz, C.c_int]
            f(_ptr(buf), *dims, int(inverse))
        else:
            f = self.lib.so_idwt3d if inverse else self.lib.so_dwt3d
            f.restype = None; f.argtypes = [vp, sz, sz, sz]
            f(_ptr(buf), *dims)
        return buf

    def dwt2d(self, vals, dims, inverse=False):
        buf = np.ascontiguousarray(vals, dtype=np.float64).copy()
        if self.is_ref:
            f = self.lib.ref_dwt2d; f.restype = None; f.argtypes = [vp, sz, sz, C.c_int]
            f(_ptr(buf), dims[0], dims[1], int(inverse))
        else:
            f = self.lib.so_idwt2d if inverse else self.lib.so_dwt2d
            f.restype = None; f.argtypes = [vp, sz, sz]
            f(_ptr(buf), dims[0], dims[1])
        return buf

    def speck3d_encode(self, mags, signs, dims, width=4, budget_bits=0):
        mags = np.ascontiguousarray(mags, dtype=np.uint64)
        signs = np.ascontiguousarray(signs, dtype=np.uint8)
        cap = mags.size * 10 + 64
        out = np.zeros(cap, dtype=np.uint8)
        if self.is_ref:
            f = self.lib.ref_speck3d_encode; f.restype = sz
            f.argtypes = [vp, vp, sz, sz, sz, C.c_int, sz, vp, sz]
            n = f(_ptr(mags), _ptr(signs), *dims, width, budget_bits, _ptr(out), cap)
        else:
            f = self.lib.so_speck3d_encode; f.restype = sz
            f.argtypes = [vp, vp, sz, sz, sz, sz, vp, sz]
            n = f(_ptr(mags), _ptr(signs), *dims, budget_bits, _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def speck3d_decode(self, stream, dims):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        n = dims[0] * dims[1] * dims[2]
        mags = np.zeros(n, dtype=np.uint64)
        signs = np.zeros(n, dtype=np.uint8)
        f = self._f("speck3d_decode"); f.restype = None; f.argtypes = [vp, sz, sz, sz, sz, vp, vp]
        f(_ptr(stream), stream.size, *dims, _ptr(mags), _ptr(signs))
        return mags, signs

    def speck2d_encode(self, mags, signs, dims, width=4, budget_bits=0):
        mags = np.ascontiguousarray(mags, dtype=np.uint64)
        signs = np.ascontiguousarray(signs, dtype=np.uint8)
        cap = mags.size * 10 + 64
        out = np.zeros(cap, dtype=np.uint8)
        if self.is_ref:
            f = self.lib.ref_speck2d_encode; f.restype = sz
            f.argtypes = [vp, vp, sz, sz, C.c_int, sz, vp, sz]
            n = f(_ptr(mags), _ptr(signs), dims[0], dims[1], width, budget_bits, _ptr(out), cap)
        else:
            f = self.lib.so_speck2d_encode; f.restype = sz
            f.argtypes = [vp, vp, sz, sz, sz, vp, sz]
            n = f(_ptr(mags), _ptr(signs), dims[0], dims[1], budget_bits, _ptr(out), cap)
        return out[:n].copy()

    def speck2d_decode(self, stream, dims):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        n = dims[0] * dims[1]
        mags = np.zeros(n, dtype=np.uint64)
        signs = np.zeros(n, dtype=np.uint8)
        f = self._f("speck2d_decode"); f.restype = None; f.argtypes = [vp, sz, sz, sz, vp, vp]
        f(_ptr(stream), stream.size, dims[0], dims[1], _ptr(mags), _ptr(signs))
        return mags, signs

    def outlier_encode(self, pos, err, total_len, tol):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        cap = total_len // 2 + pos.size * 16 + 64
        out = np.zeros(cap, dtype=np.uint8)
        f = self._f("outlier_encode"); f.restype = sz
        f.argtypes = [vp, vp, sz, sz, C.c_double, vp, sz]
        n = f(_ptr(pos), _ptr(err), pos.size, total_len, tol, _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def outlier_decode(self, stream, total_len, tol):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self._f("outlier_decode"); f.restype = sz
        f.argtypes = [vp, sz, sz, C.c_double, vp, vp, sz]
        n = f(_ptr(stream), stream.size, total_len, tol, None, None, 0)
        pos = np.zeros(n, dtype=np.uint64)
        err = np.zeros(n, dtype=np.float64)
        f(_ptr(stream), stream.size, total_len, tol, _ptr(pos), _ptr(err), n)
        return pos, err

    # ---- C API ----
    def _capi(self, name):
        return getattr(self.lib, ("sperr_" if self.is_ref else "so_") + name)

    def comp_3d(self, vol, dims, chunks, mode, quality, nthreads=0):
        vol = np.ascontiguousarray(vol)
        assert vol.dtype in (np.float32, np.float64)
        f = self._capi("comp_3d"); f.restype = C.c_int
        f.argtypes = [vp, C.c_int] + [sz] * 6 + [C.c_int, C.c_double, sz, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None); n = sz(0)
        rc = f(_ptr(vol), int(vol.dtype == np.float32), *dims, *chunks, mode, quality, nthreads,
               C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        libc_free(dst)
        return 0, out

    def decomp_3d(self, stream, output_float=True, nthreads=0):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self._capi("decomp_3d"); f.restype = C.c_int
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
        libc_free(dst)
        return 0, out, (dx.value, dy.value, dz.value)

    def trunc_3d(self, stream, pct):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self._capi("trunc_3d"); f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_uint, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None); n = sz(0)
        rc = f(_ptr(stream), stream.size, pct, C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        libc_free(dst)
        return 0, out

    def decomp_3d_multires(self, stream):
        """reference only (oracle/ref_shim.cpp over sperr::SPERR3D_OMP_D): -> (rc, full, dims,
        [coarse volumes, coarsest first], [their dims]), all fp64"""
        assert self.is_ref
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self.lib.ref_decomp_3d_multires; f.restype = C.c_int
        f.argtypes = [vp, sz, C.POINTER(sz), vp, vp, C.POINTER(vp), vp]
        nl = sz(0); ld = np.zeros(24, dtype=np.uint64); lv = (vp * 8)(); full = vp(None)
        d3 = np.zeros(3, dtype=np.uint64)
        rc = f(_ptr(stream), stream.size, C.byref(nl), _ptr(ld), C.cast(lv, vp), C.byref(full), _ptr(d3))
        if rc != 0:
            return rc, None, None, None, None
        dims = tuple(int(x) for x in d3)
        vol = np.frombuffer(C.string_at(full.value, dims[0] * dims[1] * dims[2] * 8), dtype=np.float64).copy()
        libc_free(full)
        levels, ldims = [], []
        for h in range(nl.value):
            d = tuple(int(x) for x in ld[3 * h:3 * h + 3])
            levels.append(np.frombuffer(C.string_at(lv[h], d[0] * d[1] * d[2] * 8), dtype=np.float64).copy())
            ldims.append(d)
            libc_free(lv[h])
        return 0, vol, dims, levels, ldims

    def comp_2d(self, img, dims, mode, quality, header=False):
        img = np.ascontiguousarray(img)
        f = self._capi("comp_2d"); f.restype = C.c_int
        f.argtypes = [vp, C.c_int, sz, sz, C.c_int, C.c_double, C.c_int, C.POINTER(vp), C.POINTER(sz)]
        dst = vp(None); n = sz(0)
        rc = f(_ptr(img), int(img.dtype == np.float32), dims[0], dims[1], mode, quality,
               int(header), C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.frombuffer(C.string_at(dst.value, n.value), dtype=np.uint8).copy()
        libc_free(dst)
        return 0, out

    def decomp_2d(self, stream, dims, output_float=True):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        f = self._capi("decomp_2d"); f.restype = C.c_int
        f.argtypes = [vp, sz, C.c_int, sz, sz, C.POINTER(vp)]
        dst = vp(None)
        rc = f(_ptr(stream), stream.size, int(output_float), dims[0], dims[1], C.byref(dst))
        if rc != 0:
            return rc, None
        dt = np.float32 if output_float else np.float64
        n = dims[0] * dims[1]
        out = np.frombuffer(C.string_at(dst.value, n * np.dtype(dt).itemsize), dtype=dt).copy()
        libc_free(dst)
        return 0, out


_libc = C.CDLL(None)
_libc.free.argtypes = [vp]
_libc.free.restype = None


def libc_free(p):
    _libc.free(p)


def oracle():
    return Coder(load_oracle(), "so_")


def ref():
    lib = load_ref()
    return Coder(lib, "ref_") if lib is not None else None


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------

def synthetic_field(dims, seed=1234, dtype=np.float32, modes=48, origin=(0, 0, 0)):
    """Smooth turbulence-like field: sum of separable Fourier modes with
    |k|^(-5/6) amplitudes, k in [1, 32] per 512 samples. dims = (nx, ny, nz), x fastest."""
    rng = np.random.default_rng(seed)
    k = rng.uniform(1.0, 32.0, size=(modes, 3))
    ph = rng.uniform(0.0, 2.0 * np.pi, size=(modes, 3))
    amp = np.linalg.norm(k, axis=1) ** (-5.0 / 6.0) * rng.standard_normal(modes)
    nx, ny, nz = dims
    x = (np.arange(nx) + origin[0]) * (2.0 * np.pi / 512.0)
    y = (np.arange(ny) + origin[1]) * (2.0 * np.pi / 512.0)
    z = (np.arange(nz) + origin[2]) * (2.0 * np.pi / 512.0)
    out = np.zeros((nz, ny, nx), dtype=np.float64)
    for m in range(modes):
        sx = np.sin(k[m, 0] * x + ph[m, 0])
        sy = np.sin(k[m, 1] * y + ph[m, 1])
        szz = np.sin(k[m, 2] * z + ph[m, 2])
        out += amp[m] * szz[:, None, None] * sy[None, :, None] * sx[None, None, :]
    return out.astype(dtype).reshape(-1)


def load_test_data(name, dtype=np.float32):
    """Fixtures committed under tests/golden/ (small crops) or, when present, the reference's
    own test_data directory (only in the build container)."""
    for d in (os.path.join(ROOT, "tests", "golden"), "/root/reference/test_data"):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return np.fromfile(p, dtype=dtype)
    return None
