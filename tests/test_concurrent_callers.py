"""The reference C API is re-entrant (every call builds its own encoder); here the state that
survives a call is kept per device and one mutex per device serialises the calls that target it
("Threading"). Concurrent callers must therefore be SAFE: every thread gets the
oracle's bytes, whatever the interleaving. Emulated kernels in the container, the CUDA library on a
B200."""
import threading

import numpy as np
import pytest

import gpulib
import refs


def _hammer(lib, oracle, nthreads=4, rounds=3):
    work = [
        ((32, 32, 32), (16, 16, 16), 3, 1e-3, 5),
        ((48, 40, 24), (48, 40, 24), 1, 3.0, 6),
        ((64, 32, 16), (32, 32, 16), 2, 70.0, 7),
        ((40, 24, 17), (20, 24, 17), 3, 1e-2, 8),
    ]
    expect = []
    for dims, chunks, mode, q, seed in work:
        v = refs.synthetic_field(dims, seed=seed)
        rc, s = oracle.comp_3d(v, dims, chunks, mode, q)
        assert rc == 0
        rc, d, _ = oracle.decomp_3d(s, True)
        assert rc == 0
        expect.append((v, s, d))
    errors = []

    def run(t):
        try:
            for r in range(rounds):
                i = (t + r) % len(work)
                dims, chunks, mode, q, _ = work[i]
                v, s, d = expect[i]
                rc, got = lib.comp_3d(v, dims, chunks, mode, q)
                if rc != 0 or not np.array_equal(got, s):
                    errors.append("thread %d round %d: stream differs (rc %d)" % (t, r, rc))
                rc, dec, dd = lib.decomp_3d(s, True)
                if rc != 0 or not np.array_equal(dec.view(np.uint32), d.view(np.uint32)):
                    errors.append("thread %d round %d: decoded values differ (rc %d)" % (t, r, rc))
        except Exception as e:   # noqa: BLE001 -- reported below, in the test's own thread
            errors.append("thread %d: %r" % (t, e))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_concurrent_callers_emulated(oracle):
    _hammer(gpulib.load("emul"), oracle)


@pytest.mark.gpu
def test_concurrent_callers_gpu(oracle):
    _hammer(gpulib.load("cuda"), oracle, nthreads=4, rounds=4)
