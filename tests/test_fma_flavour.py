"""The optional FMA arithmetic flavour: with sperr_b200_set_fma_flavour(1)
streams and decoded values equal those of the reference as its own build system compiles it on x86
(-mfma, -ffp-contract=fast: oracle/_ref/libsperr_ref_fma.so), where the default STRICT flavour equals
the -ffp-contract=off build. CPU part: the oracle's FMA restatement against that library, and the
kernels under the emulator; gpu part: the CUDA library against the oracle's FMA restatement."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import gpulib
import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FMA = os.path.join(ROOT, "oracle", "_ref", "libsperr_ref_fma.so")

# (fixture or synthetic dims, volume dims, chunk dims, mode, quality): fixed-rate streams are where
# the two flavours differ in bytes; PWE / PSNR streams are equal but decoded values are not
CASES = [
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 1, 4.0),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 1, 20.0),
    ("vorticity.128_128_41", (128, 128, 41), (128, 128, 41), 2, 100.0),   # wavelet-packet chunk
    ("wmag17.float", (17, 17, 17), (8, 8, 8), 1, 6.0),
    ((96, 80, 72), (96, 80, 72), (96, 80, 72), 1, 3.0),
    ((64, 64, 64), (64, 64, 64), (32, 32, 32), 3, 1e-4),
]


def _data(c):
    return refs.load_test_data(c[0]) if isinstance(c[0], str) else refs.synthetic_field(c[0], seed=5)


class FmaOracle:
    """the oracle with its lifting steps in the FMA flavour for the duration of a with-block"""

    def __init__(self, oracle):
        self.o = oracle
        self.o.lib.so_set_fma_flavour.argtypes = [C.c_int]

    def __enter__(self):
        self.o.lib.so_set_fma_flavour(1)
        return self.o

    def __exit__(self, *a):
        self.o.lib.so_set_fma_flavour(0)


def test_oracle_fma_equals_stock_reference_build(oracle):
    if not os.path.exists(REF_FMA):
        pytest.skip("oracle/_ref/libsperr_ref_fma.so not built (needs /root/reference)")
    ref = refs.Coder(C.CDLL(REF_FMA), "ref_")
    differs = 0
    for c in CASES:
        v = _data(c)
        rc, exp = ref.comp_3d(v, c[1], c[2], c[3], c[4])
        rc0, strict = oracle.comp_3d(v, c[1], c[2], c[3], c[4])
        with FmaOracle(oracle) as o:
            rc2, got = o.comp_3d(v, c[1], c[2], c[3], c[4])
            assert rc == rc2 == 0 and np.array_equal(got, exp), c
            for of in (True, False):
                _, a, _ = o.decomp_3d(exp, of)
                _, b, _ = ref.decomp_3d(exp, of)
                it = np.uint32 if of else np.uint64
                assert np.array_equal(a.view(it), b.view(it)), c
        differs += int(strict.size != exp.size or not np.array_equal(strict, exp))
    assert differs >= 3   # the flavours are really different arithmetic (fixed-rate streams)


def _check_lib(lib, oracle):
    lib.set_fma_flavour(1)
    try:
        with FmaOracle(oracle) as o:
            for c in CASES:
                v = _data(c)
                rc, got = lib.comp_3d(v, c[1], c[2], c[3], c[4])
                rc2, exp = o.comp_3d(v, c[1], c[2], c[3], c[4])
                assert rc == rc2 == 0 and np.array_equal(got, exp), c
                cases.check_decomp3d(lib, o, exp, True)
                cases.check_decomp3d(lib, o, exp, False)
    finally:
        lib.set_fma_flavour(0)
    # back to STRICT: the default flavour is untouched
    c = CASES[0]
    rc, got = lib.comp_3d(_data(c), c[1], c[2], c[3], c[4])
    rc2, exp = oracle.comp_3d(_data(c), c[1], c[2], c[3], c[4])
    assert np.array_equal(got, exp)


def test_fma_flavour_emulated(oracle):
    _check_lib(gpulib.load("emul"), oracle)


@pytest.mark.gpu
def test_fma_flavour_gpu(oracle):
    _check_lib(gpulib.load("cuda"), oracle)
