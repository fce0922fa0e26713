"""include/sperr_b200.hpp (mirrors of sperr::SPERR3D_OMP_C / SPERR3D_OMP_D): compiles with g++
against the C ABI, keeps the reference's return codes (host-only checks, no GPU needed), and -- on a
GPU -- produces the oracle's bytes and bits through the reference's sperr3d flow."""
import os
import subprocess

import numpy as np
import pytest

import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "classes_main")
SO_DIR = os.path.join(ROOT, "sperr_b200")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(os.path.join(SO_DIR, "libsperr_b200.so")):
        from sperr_b200 import build
        build.build()
    src = os.path.join(ROOT, "tests", "cpp", "classes_main.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", SO_DIR, "-lsperr_b200", "-Wl,-rpath," + SO_DIR], check=True)
    return EXE


def test_class_return_codes(exe):
    r = subprocess.run([exe, "hostonly"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout


@pytest.mark.gpu
def test_class_flow_matches_oracle(exe, oracle, tmp_path):
    v = refs.load_test_data("vorticity.128_128_41")
    fin = tmp_path / "in.f32"
    v.astype(np.float32).tofile(fin)
    for mode, q in ((3, 1e-5), (2, 90.0), (1, 3.0)):
        fs, fd = tmp_path / "s.bin", tmp_path / "d.f64"
        r = subprocess.run([exe, "run", str(fin), "128", "128", "41", "64", "64", "41", str(mode), repr(q),
                            str(fs), str(fd)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        rc, exp = oracle.comp_3d(v, (128, 128, 41), (64, 64, 41), mode, q)
        got = np.fromfile(fs, dtype=np.uint8)
        assert np.array_equal(got, exp)
        rc, dexp, dims = oracle.decomp_3d(exp, False)
        assert np.array_equal(np.fromfile(fd, dtype=np.uint64), dexp.view(np.uint64))


def test_stream_tools_mirror(exe, oracle, tmp_path):
    """SPERR3D_Stream_Tools mirror (get_header_len, get_stream_header, progressive_read,
    progressive_truncate; include/SPERR3D_Stream_Tools.h:31-46) against the unmodified reference's
    class (oracle/ref_shim.cpp) where it is built, and against the oracle's sperr_trunc_3d."""
    import ctypes as C
    v = refs.load_test_data("vorticity.128_128_41")
    ref = refs.ref()
    for chunks in ((64, 64, 41), (128, 128, 41)):
        rc, s = oracle.comp_3d(v, (128, 128, 41), chunks, 3, 1e-5)
        assert rc == 0
        fs = tmp_path / "c.sperr"
        s.tofile(fs)
        for pct in (10, 55, 100):
            prefix = tmp_path / ("t%d" % pct)
            r = subprocess.run([exe, "tools", str(fs), str(pct), str(prefix)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr + r.stdout
            rc, exp = oracle.trunc_3d(s, pct)
            assert rc == 0
            got_read = np.fromfile(str(prefix) + ".read", dtype=np.uint8)
            got_trunc = np.fromfile(str(prefix) + ".trunc", dtype=np.uint8)
            assert np.array_equal(got_read, exp) and np.array_equal(got_trunc, exp)
            fields = [int(x) for x in open(str(prefix) + ".hdr").read().split()]
            nch = 4 if chunks[0] == 64 else 1
            assert fields[:5] == [0, 0, 1, 1, int(nch > 1)] and fields[5:8] == [128, 128, 41]
            assert fields[8:11] == list(chunks) and fields[11] == (20 if nch > 1 else 14) + 4 * nch
            assert fields[12] == s.size and len(fields) == 13 + 2 * nch
            if ref is not None and hasattr(ref.lib, "ref_tools_stream_header"):
                L = ref.lib
                L.ref_tools_header_len.restype = C.c_size_t
                L.ref_tools_header_len.argtypes = [C.c_void_p]
                assert L.ref_tools_header_len(s.ctypes.data_as(C.c_void_p)) == fields[11]
                f13 = (C.c_size_t * 13)()
                offs = (C.c_size_t * 64)()
                L.ref_tools_stream_header.restype = C.c_size_t
                L.ref_tools_stream_header.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
                n = L.ref_tools_stream_header(s.ctypes.data_as(C.c_void_p), f13, offs, 64)
                assert list(f13) == fields[:13] and list(offs)[:n] == fields[13:]
                out = np.zeros(s.size + 64, dtype=np.uint8)
                L.ref_tools_progressive_read.restype = C.c_size_t
                L.ref_tools_progressive_read.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_size_t]
                m = L.ref_tools_progressive_read(str(fs).encode(), pct, out.ctypes.data_as(C.c_void_p), out.size)
                assert np.array_equal(out[:m], got_read), "progressive_read differs from the reference class"
