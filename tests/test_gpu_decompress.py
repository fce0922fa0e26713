"""GPU parity: sperr_decomp_3d of libsperr_b200.so must return values bit-identical to the
oracle's decoder for the same container, in every mode, for float and double output."""
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
def test_decomp3d_bits(lib, oracle, case):
    name, dims, chunks, mode, q = case
    v = refs.load_test_data(name)
    rc, stream = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0
    cases.check_decomp3d(lib, oracle, stream, True)
    cases.check_decomp3d(lib, oracle, stream, False)


def test_roundtrip_own_stream_pwe_bound(lib, oracle):
    # compress and decompress with the CUDA library only; the PWE bound must hold on every value
    v = refs.load_test_data("vorticity.128_128_41")
    for tol in (1e-5, 1.5e-7, 1e-2):
        rc, stream = lib.comp_3d(v, (128, 128, 41), (64, 64, 41), 3, tol)
        assert rc == 0
        rc, out, dims = lib.decomp_3d(stream, False)
        assert rc == 0 and dims == (128, 128, 41)
        assert np.max(np.abs(out - v.astype(np.float64))) <= tol


def test_decomp3d_synthetic_256(lib, oracle):
    v = refs.synthetic_field((256, 256, 256))
    for mode, q in ((3, 1e-3), (1, 1.0)):
        rc, stream = lib.comp_3d(v, (256, 256, 256), (256, 256, 256), mode, q)
        assert rc == 0
        cases.check_decomp3d(lib, oracle, stream, True)


def test_decomp3d_ragged_multichunk(lib, oracle):
    v = refs.synthetic_field((200, 150, 130), seed=7)
    for mode, q in ((3, 1e-3), (2, 70.0), (1, 1.5)):
        rc, stream = oracle.comp_3d(v, (200, 150, 130), (64, 64, 64), mode, q)
        assert rc == 0
        cases.check_decomp3d(lib, oracle, stream, True)


def test_decomp3d_rejects_bad_streams(lib, oracle):
    v = refs.load_test_data("wmag17.float")
    rc, stream = oracle.comp_3d(v, (17, 17, 17), (17, 17, 17), 3, 0.3)
    bad = stream.copy()
    bad[0] = 9                       # version byte (src/SPERR3D_OMP_D.cpp:33)
    assert lib.decomp_3d(bad)[0] == -1
    assert lib.decomp_3d(stream[:-3])[0] == -1   # length mismatch
    bad = stream.copy()
    bad[1] &= 0xBF                   # not a 3D stream
    assert lib.decomp_3d(bad)[0] == -1


@pytest.mark.parametrize("case", cases.SYN_GPU, ids=cases.syn_id)
def test_synthetic_roundtrip(lib, oracle, case):
    cases.check_syn_roundtrip(lib, oracle, case)


def test_host_api_batched_overlap_path_512(lib):
    # >= 256 MB and two z-slabs of chunks: sperr_comp_3d uploads slab groups on a copy stream while
    # the coder runs (pinned source), sperr_decomp_3d copies finished groups out while it decodes.
    # Both must give what the single-batch device-pointer path gives.
    import os
    import torch
    import sperr_b200
    os.environ["SPERR_B200_OVERLAP_MIN_CHUNKS"] = "4"   # default 256: 8 chunks would not be split
    L = sperr_b200.load()
    dims = (512, 512, 512)
    v = refs.synthetic_field(dims, seed=9)
    pinned = torch.from_numpy(v).pin_memory()
    rc, s_host = L.compress_3d(pinned.numpy(), dims, (256, 256, 256), 3, 1e-3, copy=False)
    assert rc == 0
    d_vol = pinned.cuda()
    rc, s_dev = L.compress_3d_dev(d_vol.data_ptr(), True, dims, (256, 256, 256), 3, 1e-3)
    assert rc == 0 and np.array_equal(s_host, s_dev)
    rc, out_host, d = L.decompress_3d(s_host, True, copy=False)
    assert rc == 0 and d == dims
    d_out = torch.empty(v.size, dtype=torch.float32, device="cuda")
    rc, d2 = L.decompress_3d_dev(s_dev, 0, d_out.data_ptr(), True)
    assert rc == 0 and d2 == dims
    assert np.array_equal(out_host.view(np.uint32), d_out.cpu().numpy().view(np.uint32))
    assert np.max(np.abs(out_host.astype(np.float64) - v.astype(np.float64))) <= 1e-3 + 1.2e-7
    # the same from PAGEABLE memory: a helper thread takes the slab groups through the pinned ring
    # while the coder works on the groups that have arrived
    rc, s_page = L.compress_3d(v, dims, (256, 256, 256), 3, 1e-3, copy=False)
    assert rc == 0 and np.array_equal(s_page, s_dev)
    os.environ["SPERR_B200_NO_PAGEABLE_OVERLAP"] = "1"
    rc, s_page2 = L.compress_3d(v, dims, (256, 256, 256), 3, 1e-3, copy=False)
    del os.environ["SPERR_B200_NO_PAGEABLE_OVERLAP"]
    assert rc == 0 and np.array_equal(s_page2, s_dev)
    del os.environ["SPERR_B200_OVERLAP_MIN_CHUNKS"]


@pytest.mark.parametrize("pct", [5, 30, 70])
def test_decode_truncated_container(lib, oracle, pct):
    """sperr_trunc_3d output (progressive access) decodes to the oracle's bits: truncated SPECK
    streams, outlier streams dropped when incomplete (src/SPECK_FLT.cpp:92-106)"""
    import refs
    v = refs.load_test_data("vorticity.128_128_41")
    rc, s = oracle.comp_3d(v, (128, 128, 41), (64, 64, 41), 3, 1e-5)
    assert rc == 0
    rc, t = lib.trunc_3d(s, pct)
    rc2, t2 = oracle.trunc_3d(s, pct)
    assert rc == rc2 == 0 and np.array_equal(t, t2)
    cases.check_decomp3d(lib, oracle, t, True)
