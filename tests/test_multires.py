"""Multi-resolution decoding (sperr_b200_decomp_3d_multires, the class mirror's decompress(p, true)):
coarse levels bit-identical to what the unmodified reference class returns (oracle/ref_shim.cpp over
sperr::SPERR3D_OMP_D). Emulated here, on the GPU in the gpu-marked test."""
import numpy as np
import pytest

import cases
import gpulib
import refs


def _ref():
    r = refs.ref()
    if r is None or not hasattr(r.lib, "ref_decomp_3d_multires"):
        pytest.skip("reference library with the multi-resolution shim not built")
    return r


@pytest.mark.parametrize("case", cases.MULTIRES_SMALL, ids=lambda c: "%s-%s" % (str(c[0])[:8], "x".join(map(str, c[2]))))
def test_multires_emulated(oracle, case):
    n = cases.check_multires(gpulib.load("emul"), _ref(), oracle, case)
    assert n == {(64, 64, 41): 3, (16, 16, 16): 1, (32, 32, 32): 2, (8, 8, 8): 0}[case[2]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.MULTIRES_SMALL + [((256, 256, 128), (256, 256, 128), (128, 128, 64), 3, 1e-3)],
                         ids=lambda c: "%s-%s" % (str(c[0])[:8], "x".join(map(str, c[2]))))
def test_multires_gpu(oracle, case):
    cases.check_multires(gpulib.load("cuda"), _ref(), oracle, case)
