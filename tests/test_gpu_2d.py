"""Parity gate of the 2D slice path on a B200: sperr_comp_2d / sperr_decomp_2d and the batched
entry points, through the C ABI of libsperr_b200.so, against the oracle (bytes and bits)."""
import numpy as np
import pytest

import cases
import gpulib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("cuda")


@pytest.mark.parametrize("case", cases.SLICE_GPU, ids=cases.slice_id)
def test_slice_bytes_and_bits(lib, oracle, case):
    cases.check_slice(lib, oracle, case)


def test_speck2d_stage(lib, oracle):
    rng = np.random.default_rng(11)
    for dims in ((16, 16), (37, 21), (9, 64), (100, 3), (300, 200)):
        n = dims[0] * dims[1]
        mags = (rng.standard_exponential(n) * 40).astype(np.uint64)
        mags[rng.random(n) < 0.5] = 0
        signs = (rng.random(n) < 0.5).astype(np.uint8)
        for budget in (0, 8 * (n // 3)):
            got = lib.stage_speck2d_encode(mags, signs, dims, budget)
            exp = oracle.speck2d_encode(mags, signs, dims, budget_bits=budget)
            assert np.array_equal(got, exp), (dims, budget)


def test_slice_batch(lib, oracle):
    """batched entry points: stream s equals sperr_comp_2d of slice s; constant slice included"""
    dims, ns = (256, 192), 12
    imgs = cases.slice_field(dims, np.float32, nslices=ns)
    imgs[3] = 2.5
    for mode, q in ((3, 1e-3), (2, 80.0), (1, 2.0)):
        rc, got = lib.comp_2d_batch(imgs, dims, mode, q, True)
        assert rc == 0 and len(got) == ns
        for s in range(ns):
            rc2, exp = oracle.comp_2d(imgs[s], dims, mode, q, True)
            assert rc2 == 0 and np.array_equal(got[s], exp), (mode, s)
    rc, streams = lib.comp_2d_batch(imgs, dims, 3, 1e-3, False)
    rc, dec = lib.decomp_2d_batch(streams, dims, True)
    assert rc == 0
    for s in range(ns):
        rc2, dexp = oracle.decomp_2d(streams[s], dims, True)
        assert np.array_equal(dec[s].ravel().view(np.uint32), dexp.view(np.uint32)), s
    # PWE bound holds on every slice
    assert np.max(np.abs(dec.astype(np.float64) - imgs.astype(np.float64))) <= 1e-3 * (1 + 1e-6)


def test_2d_api_errors(lib):
    img = np.zeros((8, 8), dtype=np.float32)
    assert lib.comp_2d(img, (8, 8), 3, 0.0)[0] == 2
    assert lib.comp_2d(img, (8, 8), 7, 1.0)[0] == 2
    assert lib.decomp_2d(np.zeros(3, dtype=np.uint8), (8, 8))[0] == -1
