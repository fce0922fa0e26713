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
