"""Several GPUs behind one call of the reference's C API (csrc/capi_multi.cu, SPERR_B200_DEVICES): the
container and the decoded values are the ones a single GPU produces, which are the oracle's. Skipped
on a one-GPU box (`gpurun --gpus 2`)."""
import os

import numpy as np
import pytest

import gpulib
import refs

pytestmark = pytest.mark.gpu


def test_multi_device_c_api_equals_single_device(oracle, monkeypatch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    lib = gpulib.load("cuda")
    for vol, chunk, mode, q in (((128, 96, 150), (64, 64, 48), 3, 1e-3), ((64, 64, 64), (32, 32, 32), 1, 3.0),
                                ((256, 256, 512), (256, 256, 256), 3, 1e-3)):
        v = refs.synthetic_field(vol, seed=5)
        monkeypatch.delenv("SPERR_B200_DEVICES", raising=False)
        rc, one = lib.comp_3d(v, vol, chunk, mode, q)
        assert rc == 0
        monkeypatch.setenv("SPERR_B200_DEVICES", "all")
        rc, many = lib.comp_3d(v, vol, chunk, mode, q)
        assert rc == 0
        assert np.array_equal(one, many), "multi-GPU container differs from the one-GPU container"
        if vol[0] * vol[1] * vol[2] <= 128 * 96 * 150:
            rc, exp = oracle.comp_3d(v, vol, chunk, mode, q)
            assert rc == 0 and np.array_equal(many, exp)
        rc, dm, dims = lib.decomp_3d(many, True)
        assert rc == 0 and dims == vol
        monkeypatch.delenv("SPERR_B200_DEVICES", raising=False)
        rc, d1, dims = lib.decomp_3d(one, True)
        assert rc == 0 and np.array_equal(dm.view(np.uint32), d1.view(np.uint32))
