"""Pins the C restatement (oracle/sperr_oracle.c) against the unmodified reference library
(oracle/_ref/libsperr_ref.so, STRICT flavour), stage by stage and end to end. CPU only."""
import numpy as np
import pytest

import refs

VORT = (128, 128, 41)


def _quantise(vals, q):
    ll = np.rint(vals / q).astype(np.int64)
    return np.abs(ll).astype(np.uint64), (ll >= 0).astype(np.uint8)


@pytest.mark.parametrize("n", [1, 2, 8, 9, 17, 41, 64, 91, 128, 256, 300, 2048, 4096])
def test_geometry_scalar(oracle, ref, n):
    assert oracle.num_of_xforms(n) == ref.num_of_xforms(n)
    assert oracle.num_of_partitions(n) == ref.num_of_partitions(n)
    for lev in range(0, 7):
        assert oracle.calc_approx_detail_len(n, lev) == ref.calc_approx_detail_len(n, lev)


@pytest.mark.parametrize("dims", [(64, 64, 64), (128, 128, 41), (64, 64, 41), (256, 256, 300),
                                  (17, 17, 17), (91, 91, 91), (4, 3, 8), (128, 128, 1), (10, 20, 30)])
def test_can_use_dyadic(oracle, ref, dims):
    assert oracle.can_use_dyadic(*dims) == ref.can_use_dyadic(*dims)


@pytest.mark.parametrize("vol,chunk", [((128, 128, 41), (64, 64, 41)), ((128, 128, 128), (64, 70, 80)),
                                       ((17, 17, 17), (8, 8, 8)), ((1024, 1024, 1024), (256, 256, 256)),
                                       ((100, 90, 7), (33, 200, 3))])
def test_chunk_volume(oracle, ref, vol, chunk):
    chunk = tuple(min(c, v) for c, v in zip(chunk, vol))
    assert np.array_equal(oracle.chunk_volume(vol, chunk), ref.chunk_volume(vol, chunk))


@pytest.mark.parametrize("dims", [(17, 17, 17), (64, 64, 41), (128, 128, 41), (40, 30, 21), (33, 65, 9)])
def test_conditioner_and_dwt3d(oracle, ref, dims):
    rng = np.random.default_rng(7)
    v = refs.synthetic_field(dims, seed=3, dtype=np.float64) + 0.01 * rng.standard_normal(np.prod(dims))
    a, ha = oracle.condition(v, dims)
    b, hb = ref.condition(v, dims)
    assert np.array_equal(ha, hb)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    fa = oracle.dwt3d(a, dims)
    fb = ref.dwt3d(b, dims)
    assert np.array_equal(fa.view(np.uint64), fb.view(np.uint64))
    ia = oracle.dwt3d(fa, dims, inverse=True)
    ib = ref.dwt3d(fb, dims, inverse=True)
    assert np.array_equal(ia.view(np.uint64), ib.view(np.uint64))


@pytest.mark.parametrize("dims", [(90, 90), (15, 15), (127, 64), (300, 17)])
def test_dwt2d(oracle, ref, dims):
    v = np.random.default_rng(5).standard_normal(dims[0] * dims[1])
    fa = oracle.dwt2d(v, dims)
    fb = ref.dwt2d(v, dims)
    assert np.array_equal(fa.view(np.uint64), fb.view(np.uint64))
    assert np.array_equal(oracle.dwt2d(fa, dims, True).view(np.uint64),
                          ref.dwt2d(fb, dims, True).view(np.uint64))


def test_constant_header(oracle, ref):
    v = np.full(32 * 20 * 16, 3.25)
    _, ha = oracle.condition(v, (32, 20, 16))
    _, hb = ref.condition(v, (32, 20, 16))
    assert np.array_equal(ha, hb) and ha[0] == 0x81


@pytest.mark.parametrize("dims,q,width", [((17, 17, 17), 0.05, 4), ((10, 20, 30), 0.5, 2),
                                          ((64, 64, 41), 1e-3, 4), ((33, 65, 9), 0.01, 4),
                                          ((4, 3, 8), 0.3, 1), ((128, 128, 41), 2e-3, 4),
                                          ((16, 16, 128), 1e-2, 4)])
def test_speck3d_stream_and_decode(oracle, ref, dims, q, width):
    n = int(np.prod(dims))
    v = np.random.default_rng(11).standard_normal(n) * np.exp(-np.arange(n) / (n / 6))
    mags, signs = _quantise(v, q)
    if width == 1:
        mags = np.minimum(mags, 255)
    if width == 2:
        mags = np.minimum(mags, 65535)
    sa = oracle.speck3d_encode(mags, signs, dims, width)
    sb = ref.speck3d_encode(mags, signs, dims, width)
    assert np.array_equal(sa, sb)
    ma, ga = oracle.speck3d_decode(sb, dims)
    mb, gb = ref.speck3d_decode(sb, dims)
    assert np.array_equal(ma, mb) and np.array_equal(ga, gb)
    assert np.array_equal(mb, mags)
    nz = mags != 0
    assert np.array_equal(gb[nz], signs[nz])


@pytest.mark.parametrize("bpp", [0.5, 2.0, 7.3])
def test_speck3d_budget_and_truncated_decode(oracle, ref, bpp):
    dims = (24, 17, 33)
    n = int(np.prod(dims))
    v = np.random.default_rng(2).standard_normal(n)
    mags, signs = _quantise(v, 1e-4)
    budget = int(bpp * n)
    sa = oracle.speck3d_encode(mags, signs, dims, 4, budget)
    sb = ref.speck3d_encode(mags, signs, dims, 4, budget)
    assert np.array_equal(sa, sb)
    for cut in (len(sb), len(sb) - 1, 9 + (len(sb) - 9) // 3, 10, 9):
        ma, ga = oracle.speck3d_decode(sb[:cut], dims)
        mb, gb = ref.speck3d_decode(sb[:cut], dims)
        assert np.array_equal(ma, mb) and np.array_equal(ga, gb)


def test_speck3d_all_zero(oracle, ref):
    dims = (9, 9, 9)
    mags = np.zeros(729, dtype=np.uint64)
    signs = np.ones(729, dtype=np.uint8)
    sa = oracle.speck3d_encode(mags, signs, dims)
    assert np.array_equal(sa, ref.speck3d_encode(mags, signs, dims)) and len(sa) == 9


@pytest.mark.parametrize("dims,q", [((90, 90), 0.01), ((15, 15), 0.1), ((127, 64), 1e-3), ((300, 17), 0.2)])
def test_speck2d(oracle, ref, dims, q):
    n = dims[0] * dims[1]
    v = np.random.default_rng(4).standard_normal(n) * np.exp(-np.arange(n) / (n / 5))
    mags, signs = _quantise(v, q)
    sa = oracle.speck2d_encode(mags, signs, dims)
    sb = ref.speck2d_encode(mags, signs, dims)
    assert np.array_equal(sa, sb)
    for cut in (len(sb), 9 + (len(sb) - 9) // 2):
        ma, ga = oracle.speck2d_decode(sb[:cut], dims)
        mb, gb = ref.speck2d_decode(sb[:cut], dims)
        assert np.array_equal(ma, mb) and np.array_equal(ga, gb)


@pytest.mark.parametrize("n_out,total,tol", [(3, 10, 0.1), (190, 10000, 1e-3), (3900, 900000, 1e-5),
                                             (50, 4097, 2.0)])
def test_outlier_coder(oracle, ref, n_out, total, tol):
    rng = np.random.default_rng(9)
    pos = np.sort(rng.choice(total, size=n_out, replace=False)).astype(np.uint64)
    err = tol * (1.0 + np.abs(rng.standard_normal(n_out)) * 3.0) * rng.choice([-1.0, 1.0], n_out)
    if n_out == 50:
        err[7] = 700.0  # exercises the width quirk (src/Outlier_Coder.cpp:89): u16 coder
    sa = oracle.outlier_encode(pos, err, total, tol)
    sb = ref.outlier_encode(pos, err, total, tol)
    assert np.array_equal(sa, sb)
    pa, ea = oracle.outlier_decode(sb, total, tol)
    pb, eb = ref.outlier_decode(sb, total, tol)
    assert np.array_equal(pa, pb) and np.array_equal(ea.view(np.uint64), eb.view(np.uint64))


CASES_3D = [
    # name, dims, chunks, mode, quality
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 3, 0.3),
    ("wmag17.float", (17, 17, 17), (8, 8, 8), 2, 100.0),
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 1, 2.0),
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 1, 40.0),   # high-precision retry path
    ("vorticity.128_128_41", VORT, VORT, 3, 1e-5),
    ("vorticity.128_128_41", VORT, (64, 64, 41), 3, 1.5e-7),
    ("vorticity.128_128_41", VORT, (64, 64, 41), 2, 88.0),
    ("vorticity.128_128_41", VORT, (64, 64, 41), 1, 4.0),
    ("vorticity.128_128_41", VORT, (64, 70, 30), 3, 6.7e-6),
    ("const32x20x16.float", (32, 20, 16), (32, 16, 16), 3, 1e-3),
]


@pytest.mark.parametrize("name,dims,chunks,mode,quality", CASES_3D)
def test_end_to_end_3d(oracle, ref, name, dims, chunks, mode, quality):
    vol = refs.load_test_data(name)
    assert vol is not None and vol.size == np.prod(dims)
    rc_a, sa = oracle.comp_3d(vol, dims, chunks, mode, quality)
    rc_b, sb = ref.comp_3d(vol, dims, chunks, mode, quality)
    assert rc_a == 0 and rc_b == 0
    assert np.array_equal(sa, sb)
    for of in (True, False):
        rca, da, dda = oracle.decomp_3d(sb, of)
        rcb, db, ddb = ref.decomp_3d(sb, of)
        assert rca == 0 and rcb == 0 and dda == ddb == dims
        assert np.array_equal(da.view(np.uint32 if of else np.uint64),
                              db.view(np.uint32 if of else np.uint64))
    if mode == 3:
        _, d, _ = ref.decomp_3d(sb, False)
        assert np.max(np.abs(d - vol.astype(np.float64))) <= quality


def test_end_to_end_3d_double_input(oracle, ref):
    dims = (40, 30, 21)
    vol = refs.synthetic_field(dims, dtype=np.float64)
    for mode, q in ((3, 1e-4), (2, 70.0), (1, 3.0)):
        _, sa = oracle.comp_3d(vol, dims, (20, 30, 21), mode, q)
        _, sb = ref.comp_3d(vol, dims, (20, 30, 21), mode, q)
        assert np.array_equal(sa, sb)
        _, da, _ = oracle.decomp_3d(sb, False)
        _, db, _ = ref.decomp_3d(sb, False)
        assert np.array_equal(da.view(np.uint64), db.view(np.uint64))


@pytest.mark.parametrize("name,dims,mode,quality", [("90x90.float", (90, 90), 3, 1e-2),
                                                    ("90x90.float", (90, 90), 2, 60.0),
                                                    ("90x90.float", (90, 90), 1, 2.5),
                                                    ("15x15.float", (15, 15), 3, 1e-3)])
def test_end_to_end_2d(oracle, ref, name, dims, mode, quality):
    img = refs.load_test_data(name)
    for header in (False, True):
        _, sa = oracle.comp_2d(img, dims, mode, quality, header)
        _, sb = ref.comp_2d(img, dims, mode, quality, header)
        assert np.array_equal(sa, sb)
    _, s = ref.comp_2d(img, dims, mode, quality, False)
    _, da = oracle.decomp_2d(s, dims)
    _, db = ref.decomp_2d(s, dims)
    assert np.array_equal(da.view(np.uint32), db.view(np.uint32))


def test_error_codes(oracle, ref):
    vol = refs.load_test_data("wmag17.float")
    for c in (oracle, ref):
        assert c.comp_3d(vol, (17, 17, 17), (17, 17, 17), 3, -1.0)[0] == 2
        assert c.comp_3d(vol, (17, 17, 17), (17, 17, 17), 7, 1.0)[0] == 2
        _, s = c.comp_3d(vol, (17, 17, 17), (17, 17, 17), 3, 0.1)
        assert c.decomp_3d(s[:-3])[0] == -1
