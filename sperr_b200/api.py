"""ctypes binding of include/sperr_b200.h section 1 (the reference-compatible C API,
/root/reference/include/SPERR_C_API.h:53-156) and section 2 (device-pointer extensions)."""
import ctypes as C
import os
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libsperr_b200.so")

MODE_BPP, MODE_PSNR, MODE_PWE = 1, 2, 3

sz = C.c_size_t
vp = C.c_void_p

_libc = C.CDLL(None)
_libc.free.argtypes = [vp]
_libc.free.restype = None


def _adopt(ptr, ctype, n, copy):
    """numpy view of a malloc'd result. copy=False hands the buffer itself to the caller (it is
    free()d when the array is collected), which is what a C caller of the API gets."""
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))
    if copy:
        a = a.copy()
        _libc.free(ptr)
        return a
    weakref.finalize(a, _libc.free, vp(ptr.value))
    return a


class Library:
    """One loaded libsperr_b200.so."""

    def __init__(self, path=SO_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                "%s is missing: build it with `python -m sperr_b200.build` "
                "(there is no CPU fallback)" % path)
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.sperr_comp_3d.restype = C.c_int
        L.sperr_comp_3d.argtypes = [vp, C.c_int] + [sz] * 6 + [C.c_int, C.c_double, sz,
                                                                C.POINTER(vp), C.POINTER(sz)]
        L.sperr_decomp_3d.restype = C.c_int
        L.sperr_decomp_3d.argtypes = [vp, sz, C.c_int, sz, C.POINTER(sz), C.POINTER(sz),
                                      C.POINTER(sz), C.POINTER(vp)]
        L.sperr_b200_comp_3d_dev.restype = C.c_int
        L.sperr_b200_comp_3d_dev.argtypes = [vp, C.c_int] + [sz] * 6 + [C.c_int, C.c_double,
                                                                         C.POINTER(vp), C.POINTER(sz)]
        L.sperr_b200_decomp_3d_dev.restype = C.c_int
        L.sperr_b200_decomp_3d_dev.argtypes = [vp, vp, sz, C.c_int, C.POINTER(sz), C.POINTER(sz),
                                               C.POINTER(sz), vp]
        L.sperr_comp_2d.restype = C.c_int
        L.sperr_comp_2d.argtypes = [vp, C.c_int, sz, sz, C.c_int, C.c_double, C.c_int, C.POINTER(vp),
                                    C.POINTER(sz)]
        L.sperr_decomp_2d.restype = C.c_int
        L.sperr_decomp_2d.argtypes = [vp, sz, C.c_int, sz, sz, C.POINTER(vp)]
        for name in ("sperr_b200_comp_2d_batch", "sperr_b200_comp_2d_batch_dev"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [vp, C.c_int, sz, sz, sz, C.c_int, C.c_double, C.c_int, C.POINTER(vp), vp]
        L.sperr_b200_decomp_2d_batch.restype = C.c_int
        L.sperr_b200_decomp_2d_batch.argtypes = [vp, vp, sz, C.c_int, sz, sz, C.POINTER(vp)]
        L.sperr_b200_decomp_2d_batch_dev.restype = C.c_int
        L.sperr_b200_decomp_2d_batch_dev.argtypes = [vp, vp, sz, C.c_int, sz, sz, vp]
        L.sperr_parse_header.restype = None
        L.sperr_parse_header.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz),
                                         C.POINTER(C.c_int)]

    # ---- host-buffer API (drop-in semantics) ----
    def compress_3d(self, vol, dims, chunks, mode, quality, nthreads=0, copy=True):
        """vol: flat float32/float64 array, x fastest. Returns (rc, uint8 stream or None)."""
        vol = np.ascontiguousarray(vol)
        if vol.dtype not in (np.float32, np.float64):
            raise TypeError("float32 or float64 input expected")
        if vol.size != int(dims[0]) * int(dims[1]) * int(dims[2]):
            raise ValueError("volume holds %d values, dims say %d" % (vol.size, int(dims[0]) * int(dims[1]) * int(dims[2])))
        dst, n = vp(None), sz(0)
        rc = self.lib.sperr_comp_3d(vol.ctypes.data_as(vp), int(vol.dtype == np.float32), *dims,
                                    *chunks, mode, quality, nthreads, C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        return 0, _adopt(dst, C.c_uint8, n.value, copy)

    def decompress_3d(self, stream, output_float=True, nthreads=0, copy=True):
        """Returns (rc, flat array or None, (dimx, dimy, dimz) or None)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        dx, dy, dz, dst = sz(0), sz(0), sz(0), vp(None)
        rc = self.lib.sperr_decomp_3d(stream.ctypes.data_as(vp), stream.size, int(output_float),
                                      nthreads, C.byref(dx), C.byref(dy), C.byref(dz), C.byref(dst))
        if rc != 0:
            return rc, None, None
        n = dx.value * dy.value * dz.value
        ct = C.c_float if output_float else C.c_double
        return 0, _adopt(dst, ct, n, copy), (dx.value, dy.value, dz.value)

    # ---- device-pointer extensions ----
    def compress_3d_dev(self, d_ptr, is_float, dims, chunks, mode, quality):
        """d_ptr: integer device address of the volume. Returns (rc, uint8 stream or None)."""
        dst, n = vp(None), sz(0)
        rc = self.lib.sperr_b200_comp_3d_dev(vp(d_ptr), int(is_float), *dims, *chunks, mode,
                                             quality, C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        out = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()
        _libc.free(dst)
        return 0, out

    def decompress_3d_dev(self, stream, d_stream_ptr, d_out_ptr, output_float=True):
        """stream: the container in host memory (uint8 array); d_stream_ptr: the same bytes in device
        memory (0 = upload); d_out_ptr: device buffer for the decoded volume. Returns (rc, dims)."""
        dx, dy, dz = sz(0), sz(0), sz(0)
        rc = self.lib.sperr_b200_decomp_3d_dev(stream.ctypes.data_as(vp), vp(d_stream_ptr or None),
                                               stream.size, int(output_float), C.byref(dx),
                                               C.byref(dy), C.byref(dz), vp(d_out_ptr))
        return rc, (dx.value, dy.value, dz.value)

    # ---- 2D slices (sperr_comp_2d / sperr_decomp_2d, SPERR_C_API.h:33-72) ----
    def compress_2d(self, img, dims, mode, quality, header=False):
        """img: flat float32/float64 slice, x fastest. Returns (rc, uint8 stream or None)."""
        img = np.ascontiguousarray(img)
        if img.dtype not in (np.float32, np.float64):
            raise TypeError("float32 or float64 input expected")
        if img.size != int(dims[0]) * int(dims[1]):
            raise ValueError("slice holds %d values, dims say %d" % (img.size, int(dims[0]) * int(dims[1])))
        dst, n = vp(None), sz(0)
        rc = self.lib.sperr_comp_2d(img.ctypes.data_as(vp), int(img.dtype == np.float32), dims[0],
                                    dims[1], mode, quality, int(header), C.byref(dst), C.byref(n))
        if rc != 0:
            return rc, None
        return 0, _adopt(dst, C.c_uint8, n.value, True)

    def decompress_2d(self, stream, dims, output_float=True):
        """stream: a slice stream WITHOUT the 10-byte header. Returns (rc, flat array or None)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        if int(dims[0]) * int(dims[1]) == 0:
            raise ValueError("empty slice")
        dst = vp(None)
        rc = self.lib.sperr_decomp_2d(stream.ctypes.data_as(vp), stream.size, int(output_float),
                                      dims[0], dims[1], C.byref(dst))
        if rc != 0:
            return rc, None
        return 0, _adopt(dst, C.c_float if output_float else C.c_double, dims[0] * dims[1], True)

    def compress_2d_batch(self, src, is_float, dims, nslices, mode, quality, header=False,
                          device=False):
        """src: host array of nslices contiguous slices, or (device=True) an integer device address.
        Returns (rc, uint8 buffer of all streams back to back, uint64 lengths)."""
        dst = vp(None)
        lens = np.zeros(nslices, dtype=np.uint64)
        if device:
            f, ptr = self.lib.sperr_b200_comp_2d_batch_dev, vp(src)
        else:
            src = np.ascontiguousarray(src)
            if src.dtype != (np.float32 if is_float else np.float64):
                raise TypeError("src dtype does not match is_float")
            if src.size != int(nslices) * int(dims[0]) * int(dims[1]):
                raise ValueError("src holds %d values, nslices * dims say %d"
                                 % (src.size, int(nslices) * int(dims[0]) * int(dims[1])))
            f, ptr = self.lib.sperr_b200_comp_2d_batch, src.ctypes.data_as(vp)
        rc = f(ptr, int(is_float), dims[0], dims[1], nslices, mode, quality, int(header),
               C.byref(dst), lens.ctypes.data_as(vp))
        if rc != 0:
            return rc, None, None
        return 0, _adopt(dst, C.c_uint8, int(lens.sum()), False), lens

    def decompress_2d_batch(self, streams, lens, dims, output_float=True, d_out_ptr=None):
        """streams: uint8 buffer of headerless slice streams back to back; lens: their lengths.
        Returns (rc, flat array) or, with d_out_ptr (device buffer), (rc, None)."""
        streams = np.ascontiguousarray(streams, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        if int(lens.sum()) > streams.size:
            raise ValueError("lens add up to %d bytes, the buffer holds %d" % (int(lens.sum()), streams.size))
        if d_out_ptr is not None:
            rc = self.lib.sperr_b200_decomp_2d_batch_dev(streams.ctypes.data_as(vp),
                                                         lens.ctypes.data_as(vp), lens.size,
                                                         int(output_float), dims[0], dims[1],
                                                         vp(d_out_ptr))
            return rc, None
        dst = vp(None)
        rc = self.lib.sperr_b200_decomp_2d_batch(streams.ctypes.data_as(vp), lens.ctypes.data_as(vp),
                                                 lens.size, int(output_float), dims[0], dims[1],
                                                 C.byref(dst))
        if rc != 0:
            return rc, None
        ct = C.c_float if output_float else C.c_double
        return 0, _adopt(dst, ct, dims[0] * dims[1] * lens.size, False)

    def parse_header(self, stream):
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        dx, dy, dz, isf = sz(0), sz(0), sz(0), C.c_int(0)
        self.lib.sperr_parse_header(stream.ctypes.data_as(vp), C.byref(dx), C.byref(dy),
                                    C.byref(dz), C.byref(isf))
        return (dx.value, dy.value, dz.value), bool(isf.value)

    def fn(self, name, restype, argtypes):
        f = getattr(self.lib, name)
        f.restype = restype
        f.argtypes = argtypes
        return f


_default = None


def load(path=SO_PATH):
    global _default
    if _default is None or _default.path != path:
        _default = Library(path)
    return _default


def compress_3d(vol, dims, chunks, mode, quality):
    return load().compress_3d(vol, dims, chunks, mode, quality)


def decompress_3d(stream, output_float=True):
    return load().decompress_3d(stream, output_float)


def parse_header(stream):
    return load().parse_header(stream)
