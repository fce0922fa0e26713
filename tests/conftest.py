import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def oracle():
    import refs
    return refs.oracle()


@pytest.fixture(scope="session")
def ref():
    import refs
    r = refs.ref()
    if r is None:
        pytest.skip("oracle/_ref/libsperr_ref.so not built (needs /root/reference)")
    return r
