"""Volumes taken through HBM one group of chunk slabs at a time (csrc/capi_stream.cu): forced here by
SPERR_B200_STREAM_MB on small volumes; the container and the decoded bits must be what the resident
path -- and the oracle -- give. Emulated kernels in the container, the CUDA library on a B200."""
import numpy as np
import pytest

import gpulib
import refs

# (dims, chunk dims, mode, quality, group limit in MB)
CASES = [
    ((64, 64, 128), (32, 32, 32), 3, 1e-3, 1),      # 4 slabs of 4 chunks, 2 slabs per group
    ((64, 48, 100), (32, 32, 32), 1, 3.0, 1),       # ragged chunks, fixed rate, 3 slabs in 2 groups
    ((128, 64, 64), (64, 64, 16), 2, 70.0, 1),      # 4 thin slabs
]


def _check(lib, oracle, case, monkeypatch):
    dims, chunks, mode, q, mb = case
    v = refs.synthetic_field(dims, seed=21)
    rc, exp = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0
    monkeypatch.setenv("SPERR_B200_STREAM_MB", str(mb))
    assert v.nbytes > mb << 20
    rc, got = lib.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0
    assert np.array_equal(got, exp), "streamed container differs from the oracle's"
    rc, dec, d = lib.decomp_3d(exp, True)
    rc2, dexp, d2 = oracle.decomp_3d(exp, True)
    assert rc == 0 and rc2 == 0 and tuple(d) == tuple(d2)
    assert np.array_equal(dec.view(np.uint32), dexp.view(np.uint32)), "streamed decode differs"
    rc, dec64, d = lib.decomp_3d(exp, False)
    rc2, dexp64, d2 = oracle.decomp_3d(exp, False)
    assert rc == 0 and np.array_equal(dec64.view(np.uint64), dexp64.view(np.uint64))
    # the resident path gives the same container
    monkeypatch.delenv("SPERR_B200_STREAM_MB")
    rc, got2 = lib.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0 and np.array_equal(got2, exp)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[0])))
def test_streamed_emulated(oracle, case, monkeypatch):
    _check(gpulib.load("emul"), oracle, case, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [((256, 256, 512), (256, 256, 128), 3, 1e-3, 40)],
                         ids=lambda c: "x".join(map(str, c[0])))
def test_streamed_gpu(oracle, case, monkeypatch):
    _check(gpulib.load("cuda"), oracle, case, monkeypatch)
