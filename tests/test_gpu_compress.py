"""GPU parity: sperr_comp_3d of libsperr_b200.so (C ABI, CUDA kernels) must produce streams
byte-identical to the oracle's (oracle/sperr_oracle.c, pinned against the reference build)."""
import numpy as np
import pytest

import cases
import gpulib
import refs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("cuda")


@pytest.mark.parametrize("case", cases.COMP3D_GPU, ids=cases.case_id)
def test_comp3d_bytes(lib, oracle, case):
    cases.check_comp3d(lib, oracle, case)


def test_comp3d_f64_input(lib, oracle):
    cases.check_comp3d(lib, oracle, cases.COMP3D_GPU[6], f64=True)


def test_comp3d_synthetic_256_pwe(lib, oracle):
    # one full-size chunk of the benchmark workload (SURVEY.md 8d synthetic field), 5 dyadic levels
    v = refs.synthetic_field((256, 256, 256))
    rc, got = lib.comp_3d(v, (256, 256, 256), (256, 256, 256), 3, 1e-3)
    rc2, exp = oracle.comp_3d(v, (256, 256, 256), (256, 256, 256), 3, 1e-3)
    assert rc == rc2 == 0
    assert np.array_equal(got, exp)


def test_comp3d_synthetic_multichunk_ragged(lib, oracle):
    v = refs.synthetic_field((200, 150, 130), seed=7)
    for mode, q in ((3, 1e-3), (2, 70.0), (1, 1.5)):
        rc, got = lib.comp_3d(v, (200, 150, 130), (64, 64, 64), mode, q)
        rc2, exp = oracle.comp_3d(v, (200, 150, 130), (64, 64, 64), mode, q)
        assert rc == rc2 == 0
        assert np.array_equal(got, exp), (mode, q)


def test_error_codes(lib):
    v = np.zeros(8, dtype=np.float32)
    assert lib.comp_3d(v, (2, 2, 2), (2, 2, 2), 3, 0.0)[0] == 2
    assert lib.comp_3d(v, (2, 2, 2), (2, 2, 2), 7, 1.0)[0] == 2


def test_stage_dwt_bits(lib, oracle):
    rng = np.random.default_rng(3)
    for dims in ((91, 91, 91), (128, 128, 41), (64, 33, 20), (256, 64, 64)):
        v = rng.standard_normal(dims[0] * dims[1] * dims[2])
        a = lib.stage_dwt(v, dims)
        b = oracle.dwt3d(v, dims)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), dims
        a2 = lib.stage_dwt(a, dims, inverse=True)
        b2 = oracle.dwt3d(b, dims, inverse=True)
        assert np.array_equal(a2.view(np.uint64), b2.view(np.uint64)), dims


def test_stage_dwt_fused_bits(lib, oracle):
    # the one-round-trip-per-level kernels (dyadic shapes): bit-identical coefficients both ways
    rng = np.random.default_rng(4)
    for dims in ((91, 91, 91), (256, 256, 256), (40, 24, 17), (128, 128, 128), (300, 280, 290)):
        v = rng.standard_normal(dims[0] * dims[1] * dims[2])
        rc, a = lib.stage_dwt_fused(v, dims)
        assert rc == 0, (dims, rc)
        b = oracle.dwt3d(v, dims)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), dims
        rc, a2 = lib.stage_dwt_fused(b, dims, inverse=True)
        b2 = oracle.dwt3d(b, dims, inverse=True)
        assert rc == 0 and np.array_equal(a2.view(np.uint64), b2.view(np.uint64)), dims
