"""The product library loads without a GPU and exports every symbol include/sperr_b200.h declares;
compute entry points fail loudly (-1) when no CUDA device is present: there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sperr_b200", "libsperr_b200.so")
HEADER = os.path.join(ROOT, "include", "sperr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sperr_b200|sperr)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        from sperr_b200 import build
        build.build()
    return C.CDLL(SO)


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert "sperr_comp_3d" in names and "sperr_decomp_3d" in names and len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    v = np.zeros(16 ** 3, dtype=np.float32)
    dst, n = C.c_void_p(None), C.c_size_t(0)
    lib.sperr_comp_3d.restype = C.c_int
    lib.sperr_comp_3d.argtypes = [C.c_void_p, C.c_int] + [C.c_size_t] * 6 + [C.c_int, C.c_double, C.c_size_t,
                                                                              C.POINTER(C.c_void_p),
                                                                              C.POINTER(C.c_size_t)]
    rc = lib.sperr_comp_3d(v.ctypes.data_as(C.c_void_p), 1, 16, 16, 16, 16, 16, 16, 3, 1e-3, 0, C.byref(dst),
                           C.byref(n))
    assert rc == -1 and not dst.value


def test_python_package_requires_the_library(tmp_path):
    from sperr_b200 import api
    with pytest.raises(RuntimeError):
        api.Library(str(tmp_path / "missing.so"))
