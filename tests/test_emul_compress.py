"""Kernel-logic checks in the build container: the CUDA sources compiled against the CPU SIMT
emulator (tests/emul, test infrastructure only) must already be byte-identical to the oracle on
small inputs. The real parity gate is test_gpu_*.py on a B200."""
import pytest

import cases
import gpulib


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("emul")


@pytest.mark.parametrize("case", cases.COMP3D_SMALL, ids=cases.case_id)
def test_comp3d_bytes_emulated(lib, oracle, case):
    cases.check_comp3d(lib, oracle, case)


def test_stage_dwt_fused_bits_emulated(lib, oracle):
    import numpy as np
    rng = np.random.default_rng(4)
    for dims in ((16, 16, 16), (91, 91, 91), (40, 24, 17), (64, 64, 64)):
        v = rng.standard_normal(dims[0] * dims[1] * dims[2])
        rc, a = lib.stage_dwt_fused(v, dims)
        assert rc == 0, (dims, rc)
        b = oracle.dwt3d(v, dims)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), dims
        rc, a2 = lib.stage_dwt_fused(b, dims, inverse=True)
        b2 = oracle.dwt3d(b, dims, inverse=True)
        assert rc == 0 and np.array_equal(a2.view(np.uint64), b2.view(np.uint64)), dims
