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
