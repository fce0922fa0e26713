"""Golden vectors and the reference's own known-answer tests, without the reference at run time.

* tests/golden/reference_fingerprints.json was written by tests/golden/make_golden.py from the
  UNMODIFIED reference (stream length + FNV-1a 64 of the stream and of the decoded values, per
  fixture and configuration). The oracle is held to it on the CPU, the CUDA library on a B200, so
  both stay pinned on a box where /root/reference and oracle/_ref do not exist. The stream lengths
  of the vorticity cases equal the ones SURVEY.md 8c recorded independently (43 596, 957 708,
  477 054, 1 166 729, 336 012, 1 679 500 bytes).
* The assertions of the reference's unit tests for this path, run on the oracle:
  speck3d_flt_unit_test.cpp:29,49 (constant field: 17-byte chunk stream), :94-146 (integer width
  1 / 2 / 4 / 8 at PSNR 40 / 50 / 190 / 210 on wmag17), :152-236 (PWE bound at 1e-5, 2.9e-9, 1e-2),
  sperr3d_omp_unit_test.cpp:84-122 (PWE 1.5e-7, 6.7e-6 with 64x64x41 chunks), :211-253 (PSNR
  windows 89.1123..89.1124 at 88 dB and 126.8866..126.8867 at 125 dB, chunks 64^3).
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases      # noqa: E402
import gpulib     # noqa: E402
import refs       # noqa: E402
import make_golden as mg   # noqa: E402

REC = json.load(open(os.path.join(HERE, "golden", "reference_fingerprints.json")))["records"]


def _check3d(coder, c):
    rec = REC[mg.key3d(c)]
    v = mg.load3d(c)
    rc, s = coder.comp_3d(v, c[1], c[2], c[3], c[4])[:2]
    assert rc == 0
    assert int(s.size) == rec["len"]
    assert cases.fnv1a64(s) == rec["stream"]
    rc, d32, dims = coder.decomp_3d(s, True)[:3]
    assert rc == 0 and tuple(dims) == tuple(c[1])
    assert cases.fnv1a64(np.ascontiguousarray(d32).tobytes()) == rec["dec_f32"]
    rc, d64, _ = coder.decomp_3d(s, False)[:3]
    assert rc == 0
    assert cases.fnv1a64(np.ascontiguousarray(d64).tobytes()) == rec["dec_f64"]


def _check2d(coder, c):
    rec = REC[mg.key2d(c)]
    img = refs.load_test_data(c[0])
    rc, s = coder.comp_2d(img, c[1], c[2], c[3], False)[:2]
    assert rc == 0
    assert int(s.size) == rec["len"] and cases.fnv1a64(s) == rec["stream"]
    rc, d32 = coder.decomp_2d(s, c[1], True)[:2]
    assert rc == 0
    assert cases.fnv1a64(np.ascontiguousarray(d32).tobytes()) == rec["dec_f32"]


def test_every_record_has_a_case():
    keys = {mg.key3d(c) for c in mg.CASES_3D} | {mg.key2d(c) for c in mg.CASES_2D}
    assert keys == set(REC)


@pytest.mark.parametrize("c", mg.CASES_3D, ids=mg.key3d)
def test_oracle_reproduces_reference_3d(oracle, c):
    _check3d(oracle, c)


@pytest.mark.parametrize("c", mg.CASES_2D, ids=mg.key2d)
def test_oracle_reproduces_reference_2d(oracle, c):
    _check2d(oracle, c)


# ---- the reference's own known answers, on the oracle -------------------------------------------

def _psnr(orig, rec):
    # sperr::calc_stats, /root/reference/src/sperr_helper.cpp:429-518: 10 log10(range^2 / mse)
    o, r = orig.astype(np.float64), rec.astype(np.float64)
    rng = o.max() - o.min()
    mse = np.mean((o - r) ** 2)
    return 10.0 * np.log10(rng * rng / mse), np.abs(o - r).max()


def _container_chunks(stream):
    """(header length, chunk lengths) of a 3D container (include/bitstream_definition.txt)."""
    s = bytes(stream)
    assert s[1] & 0x40   # 3D
    multi = bool(s[1] & 0x10)
    pos = 2 + 12 + (6 if multi else 0)
    vol = np.frombuffer(s[2:14], dtype=np.uint32)
    ch = np.frombuffer(s[14:20], dtype=np.uint16).astype(np.int64) if multi else vol.astype(np.int64)
    n = 1
    for a in range(3):
        k = int(vol[a]) // int(ch[a])
        if int(vol[a]) % int(ch[a]) > int(ch[a]) // 2:
            k += 1
        n *= max(k, 1)
    lens = np.frombuffer(s[pos:pos + 4 * n], dtype=np.uint32)
    return pos + 4 * n, lens


def test_constant_field_is_17_bytes_per_chunk(oracle):
    v = refs.load_test_data("const32x20x16.float")
    rc, s = oracle.comp_3d(v, (32, 20, 16), (16, 16, 16), 3, 1e-3)
    assert rc == 0
    hlen, lens = _container_chunks(s)
    assert list(lens) == [17] * len(lens) and s.size == hlen + 17 * len(lens)
    rc, d, _ = oracle.decomp_3d(s, True)
    assert rc == 0 and np.array_equal(d, v)


@pytest.mark.parametrize("psnr,width", [(40.0, 1), (50.0, 2), (190.0, 4), (210.0, 8)])
def test_integer_width_on_wmag17(oracle, psnr, width):
    v = refs.load_test_data("wmag17.float").astype(np.float64)
    rc, s = oracle.comp_3d(v, (17, 17, 17), (17, 17, 17), 2, psnr)
    assert rc == 0
    hlen, lens = _container_chunks(s)
    planes = int(s[hlen + 17])   # SPECK stream header: u8 bit planes (src/SPECK_INT.cpp:284-308)
    got = 1 if planes <= 8 else 2 if planes <= 16 else 4 if planes <= 32 else 8   # SPECK_FLT.cpp:64-72
    assert got == width


@pytest.mark.parametrize("chunks,tol", [((128, 128, 41), 1e-5), ((128, 128, 41), 2.9e-9), ((128, 128, 41), 1e-2),
                                        ((64, 64, 41), 1.5e-7), ((64, 64, 41), 6.7e-6)])
def test_pwe_bound_on_vorticity(oracle, chunks, tol):
    v = refs.load_test_data("vorticity.128_128_41")
    rc, s = oracle.comp_3d(v.astype(np.float64), (128, 128, 41), chunks, 3, tol)
    assert rc == 0
    rc, d, _ = oracle.decomp_3d(s, False)
    assert rc == 0
    assert np.abs(d - v.astype(np.float64)).max() <= tol


@pytest.mark.parametrize("target,lo,hi", [(88.0, 89.1123, 89.1124), (125.0, 126.8866, 126.8867)])
def test_psnr_windows_on_vorticity(oracle, target, lo, hi):
    v = refs.load_test_data("vorticity.128_128_41")
    rc, s = oracle.comp_3d(v, (128, 128, 41), (64, 64, 64), 2, target)
    assert rc == 0
    rc, d, _ = oracle.decomp_3d(s, False)
    assert rc == 0
    psnr, _ = _psnr(v, d)
    assert lo < psnr < hi


# ---- the CUDA sources under the CPU SIMT emulator (test infrastructure) against the records ------

@pytest.fixture(scope="module")
def emul_lib():
    return gpulib.load("emul")


_EMUL_3D = [c for c in mg.CASES_3D if c[:5] in cases.COMP3D_SMALL or c[5] == "f64"]


@pytest.mark.parametrize("c", _EMUL_3D, ids=mg.key3d)
def test_emulated_kernels_reproduce_reference_3d(emul_lib, c):
    _check3d(emul_lib, c)


@pytest.mark.parametrize("c", mg.CASES_2D, ids=mg.key2d)
def test_emulated_kernels_reproduce_reference_2d(emul_lib, c):
    _check2d(emul_lib, c)


# ---- the CUDA library against the same records (B200) --------------------------------------------

@pytest.fixture(scope="module")
def cuda_lib():
    return gpulib.load("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("c", mg.CASES_3D, ids=mg.key3d)
def test_cuda_reproduces_reference_3d(cuda_lib, c):
    _check3d(cuda_lib, c)


@pytest.mark.gpu
@pytest.mark.parametrize("c", mg.CASES_2D, ids=mg.key2d)
def test_cuda_reproduces_reference_2d(cuda_lib, c):
    _check2d(cuda_lib, c)
