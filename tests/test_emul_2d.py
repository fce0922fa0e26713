"""2D slice path (sperr_comp_2d / sperr_decomp_2d and the batched entry points) under the CPU SIMT
emulator: container-side logic check on small slices; the parity gate is test_gpu_2d.py."""
import numpy as np
import pytest

import cases
import gpulib


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("emul")


@pytest.mark.parametrize("case", cases.SLICE_SMALL, ids=cases.slice_id)
def test_slice_bytes_and_bits_emulated(lib, oracle, case):
    cases.check_slice(lib, oracle, case)


def test_speck2d_stage_emulated(lib, oracle):
    rng = np.random.default_rng(11)
    for dims in ((16, 16), (37, 21), (9, 64), (100, 3)):
        n = dims[0] * dims[1]
        mags = (rng.standard_exponential(n) * 40).astype(np.uint64)
        mags[rng.random(n) < 0.5] = 0
        signs = (rng.random(n) < 0.5).astype(np.uint8)
        for budget in (0, 8 * (n // 3)):
            got = lib.stage_speck2d_encode(mags, signs, dims, budget)
            exp = oracle.speck2d_encode(mags, signs, dims, budget_bits=budget)
            assert np.array_equal(got, exp), (dims, budget)


def test_slice_batch_emulated(lib, oracle):
    dims, ns = (48, 40), 5
    imgs = cases.slice_field(dims, np.float32, nslices=ns)
    imgs[3] = 2.5   # a constant slice: 17-byte stream
    for header in (False, True):
        rc, got = lib.comp_2d_batch(imgs, dims, 3, 1e-3, header)
        assert rc == 0 and len(got) == ns
        for s in range(ns):
            rc2, exp = oracle.comp_2d(imgs[s], dims, 3, 1e-3, header)
            assert rc2 == 0 and np.array_equal(got[s], exp), s
    rc, streams = lib.comp_2d_batch(imgs, dims, 3, 1e-3, False)
    rc, dec = lib.decomp_2d_batch(streams, dims, True)
    assert rc == 0
    for s in range(ns):
        rc2, dexp = oracle.decomp_2d(streams[s], dims, True)
        assert np.array_equal(dec[s].ravel().view(np.uint32), dexp.view(np.uint32)), s


def test_2d_api_errors_emulated(lib):
    img = np.zeros((8, 8), dtype=np.float32)
    assert lib.comp_2d(img, (8, 8), 3, 0.0)[0] == 2
    assert lib.comp_2d(img, (8, 8), 7, 1.0)[0] == 2
    assert lib.decomp_2d(np.zeros(3, dtype=np.uint8), (8, 8))[0] == -1
