"""BASELINE.json configurations at (or scaled towards) their full sizes, checked through properties
that do not need the oracle to code the whole volume:
  * config #2 at full size (1024^3 fp32, PWE 1e-3, 256^3 chunks): the PWE bound holds on every value,
    decoding is idempotent, the container is self-consistent, and -- because SPERR's chunks are coded
    independently (src/SPERR3D_OMP_C.cpp:94-130) -- the stream of any chunk inside the container is
    byte-identical to what the ORACLE produces for that 256^3 sub-volume alone (sampled chunks);
  * config #3 scaled (fp64, fixed rate 2 bpp, 256^3 chunks) and config #4 scaled (fp32, PSNR target,
    decompression) against the oracle on two chunks."""
import os
import sys

import numpy as np
import pytest

import cases
import gpulib
import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return gpulib.load("cuda")


def test_config2_full_size_properties(lib, oracle):
    import torch

    import bench
    import sperr_b200
    from sperr_b200 import sharded

    L = sperr_b200.load()
    n, ck, tol = 1024, 256, 1e-3
    dev = torch.device("cuda", 0)
    vol = bench.field_torch((n, n, n), (0, 0, 0), dev)
    rc, stream = L.compress_3d_dev(vol.data_ptr(), True, (n, n, n), (ck, ck, ck), 3, tol)
    assert rc == 0
    v, c, isf, hlen, lens = sharded.parse_container(L.lib, stream)
    assert v == (n, n, n) and c == (ck, ck, ck) and isf and lens.size == 64
    assert hlen + int(lens.astype(np.int64).sum()) == stream.size
    # decode twice: identical bits, PWE bound on every value
    out = torch.empty_like(vol)
    rc, dims = L.decompress_3d_dev(stream, 0, out.data_ptr(), True)
    assert rc == 0 and dims == (n, n, n)
    err = float((out.double() - vol.double()).abs().max())
    assert err <= tol + 1.2e-7, err   # the bound is enforced in fp64, the result is rounded to fp32
    out2 = torch.empty_like(vol)
    rc, _ = L.decompress_3d_dev(stream, 0, out2.data_ptr(), True)
    assert rc == 0 and torch.equal(out.view(torch.int32), out2.view(torch.int32))
    # chunk independence: sampled chunk streams equal the oracle's stream of the sub-volume alone
    offs = hlen + np.concatenate([[0], np.cumsum(lens.astype(np.int64))])
    v3 = vol.view(n, n, n)
    for k in (0, 27, 63):
        cz, cy, cx = k // 16, (k // 4) % 4, k % 4
        sub = v3[cz * ck:(cz + 1) * ck, cy * ck:(cy + 1) * ck, cx * ck:(cx + 1) * ck].contiguous().cpu().numpy()
        rc2, exp = oracle.comp_3d(sub.reshape(-1), (ck, ck, ck), (ck, ck, ck), 3, tol)
        assert rc2 == 0
        got = np.asarray(stream[offs[k]:offs[k + 1]])
        assert np.array_equal(got, exp[14 + 4:]), "chunk %d differs from the oracle" % k
    # and the decoded chunk equals the oracle's decode of that chunk stream
    rc3, dexp, d3 = oracle.decomp_3d(exp, True)
    assert rc3 == 0
    dsub = out.view(n, n, n)[3 * ck:, 3 * ck:, 3 * ck:].contiguous().cpu().numpy().reshape(-1)
    assert np.array_equal(dsub.view(np.uint32), dexp.view(np.uint32))


def test_config3_scaled_f64_fixed_rate(lib, oracle):
    dims, ck = (512, 256, 256), (256, 256, 256)
    v = refs.synthetic_field(dims, seed=11, dtype=np.float64)
    rc, got = lib.comp_3d(v, dims, ck, 1, 2.0)
    rc2, exp = oracle.comp_3d(v, dims, ck, 1, 2.0)
    assert rc == rc2 == 0
    assert np.array_equal(got, exp)
    # fixed rate: every chunk stream is 26 header bytes + 2 bits per value
    assert got.size == 20 + 8 + 2 * (26 + 2 * 256 ** 3 // 8)
    cases.check_decomp3d(lib, oracle, exp, False)


def test_config4_scaled_psnr_decompress(lib, oracle):
    dims, ck = (512, 256, 256), (256, 256, 256)
    v = refs.synthetic_field(dims, seed=12)
    rc, exp = oracle.comp_3d(v, dims, ck, 2, 80.0)
    assert rc == 0
    dec = cases.check_decomp3d(lib, oracle, exp, True)
    rng = float(v.max() - v.min())
    mse = float(np.mean((dec.astype(np.float64) - v.astype(np.float64)) ** 2))
    assert 10 * np.log10(rng * rng / mse) >= 79.9   # the quantiser aims at the target from above
