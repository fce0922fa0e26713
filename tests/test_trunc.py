"""sperr_trunc_3d (progressive truncation, /root/reference/src/SPERR3D_Stream_Tools.cpp:131-226):
host-only byte surgery, so the CUDA library's entry point is checked here without a GPU against the
oracle and (where built) the unmodified reference; decoding truncated containers is a gpu / emulator
test."""
import numpy as np
import pytest

import cases
import gpulib
import refs


def _streams(oracle):
    v = refs.load_test_data("vorticity.128_128_41")
    out = []
    for chunks, mode, q in (((64, 64, 41), 3, 1e-5), ((128, 128, 41), 2, 100.0), ((64, 70, 20), 1, 6.0)):
        rc, s = oracle.comp_3d(v, (128, 128, 41), chunks, mode, q)
        assert rc == 0
        out.append(s)
    c = refs.load_test_data("const32x20x16.float")
    rc, s = oracle.comp_3d(c, (32, 20, 16), (16, 16, 16), 3, 1e-3)   # constant chunks: 17 bytes each
    out.append(s)
    return out


@pytest.fixture(scope="module")
def cuda_lib():
    return gpulib.load("cuda")


def test_trunc_matches_oracle_and_reference(cuda_lib, oracle):
    ref = refs.ref()
    for s in _streams(oracle):
        for pct in (0, 1, 10, 37, 50, 99, 100, 250):
            rc, got = cuda_lib.trunc_3d(s, pct)
            rc2, exp = oracle.trunc_3d(s, pct)
            assert rc == rc2 == 0
            assert np.array_equal(got, exp), pct
            if ref is not None:
                rc3, r = ref.trunc_3d(s, pct)
                assert rc3 == 0 and np.array_equal(got, r), pct
            if 0 < pct < 100:
                assert got[1] & 0x80 and got.size <= s.size
            else:
                assert np.array_equal(got, s)


def test_trunc_errors(cuda_lib, oracle):
    s = _streams(oracle)[0]
    assert cuda_lib.trunc_3d(s[:10], 50)[0] == -1
    assert cuda_lib.trunc_3d(s[: s.size // 2], 90)[0] == -1   # needs bytes beyond the given length
    # a prefix that is long enough for the requested cut is accepted (single-chunk use case)
    one = _streams(oracle)[1]
    rc, a = cuda_lib.trunc_3d(one, 10)
    rc2, b = cuda_lib.trunc_3d(one[: one.size // 10 + 64 + 40], 10)
    assert rc == rc2 == 0 and np.array_equal(a, b)


def test_decode_truncated_emulated(oracle):
    """a truncated container decodes to the same bits as the oracle decodes it to"""
    lib = gpulib.load("emul")
    v = refs.load_test_data("wmag17.float")
    rc, s = oracle.comp_3d(v, (17, 17, 17), (8, 8, 8), 3, 0.05)
    for pct in (30, 70):
        rc, t = lib.trunc_3d(s, pct)
        assert rc == 0
        cases.check_decomp3d(lib, oracle, t, True)
