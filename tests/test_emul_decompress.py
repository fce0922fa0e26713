"""Decoder kernel logic under the CPU SIMT emulator (test infrastructure; see test_emul_compress)."""
import pytest

import cases
import gpulib
import refs


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("emul")


@pytest.mark.parametrize("case", cases.COMP3D_SMALL, ids=cases.case_id)
def test_decomp3d_bits_emulated(lib, oracle, case):
    name, dims, chunks, mode, q = case
    v = refs.load_test_data(name)
    rc, stream = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0
    cases.check_decomp3d(lib, oracle, stream, True)


@pytest.mark.parametrize("case", cases.SYN_SMALL, ids=cases.syn_id)
def test_synthetic_roundtrip_emulated(lib, oracle, case):
    cases.check_syn_roundtrip(lib, oracle, case)


@pytest.mark.parametrize("R", [1, 2, 8])
def test_cluster_decode_emulated(lib, oracle, R, monkeypatch):
    """Thread-block clusters (csrc/speck_dec_fast.cuh): R CTAs resolve R consecutive windows of one
    stream per round; the decoded bits do not depend on R. 128^3 chunk: about 30 (PWE) and 1000
    (4 bpp) windows of 8192 bits, and a 2^21-leaf outlier tree for the 1D walker."""
    monkeypatch.setenv("SPERR_B200_DEC_CLUSTER", str(R))
    dims = (128, 128, 128)
    v = refs.synthetic_field(dims, seed=5)
    for mode, q in ((3, 1e-3), (1, 4.0)):
        rc, s = oracle.comp_3d(v, dims, dims, mode, q)
        assert rc == 0
        cases.check_decomp3d(lib, oracle, s, True)
    # with a noise floor: a dense LIP for the cluster's split mask sweep, and an outlier tree with
    # both single-path subtrees (taken in one step by the 1D walker) and crowded ones (bit by bit)
    import numpy as np
    vn = (v + 3e-4 * np.random.default_rng(11).standard_normal(v.shape)).astype(np.float32)
    rc, s = oracle.comp_3d(vn, dims, dims, 3, 1e-3)
    assert rc == 0
    cases.check_decomp3d(lib, oracle, s, True)
