"""Shared parity cases: the same checks run against the CUDA library on a B200 (test_gpu_*.py,
marked gpu) and -- small sizes only -- against the CPU SIMT emulation of the same kernels
(test_emul_*.py), which is a container-side debugging aid and never a product path."""
import numpy as np

import refs

# (fixture, dims, chunk dims, mode, quality)
COMP3D_SMALL = [
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 3, 0.3),
    ("wmag17.float", (17, 17, 17), (8, 8, 8), 2, 100.0),       # 8 ragged chunks
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 1, 2.0),
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 1, 40.0),     # fixed-rate high-precision retry
    ("wmag16.float", (16, 16, 16), (16, 16, 16), 2, 60.0),
    ("const32x20x16.float", (32, 20, 16), (16, 16, 16), 3, 1e-3),   # constant chunks: 17-byte streams
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 3, 1e-5),
    ("vorticity.128_128_41", (128, 128, 41), (128, 128, 41), 3, 1e-5),   # wavelet-packet (non-dyadic)
]

COMP3D_GPU = COMP3D_SMALL + [
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 3, 2.9e-9),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 3, 1.5e-7),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 64), 2, 88.0),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 64), 2, 125.0),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 1, 4.0),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 1, 20.0),
    ("vorticity.128_128_41", (128, 128, 41), (64, 70, 20), 3, 6.7e-6),   # ragged
]


# synthetic fields (refs.synthetic_field, seed 5): (dims, chunk dims, mode, quality).
# Power-of-two dyadic chunks take the decoder's shift-addressed fast path, the others the table path.
SYN_SMALL = [
    ((32, 32, 32), (16, 16, 16), 3, 1e-3),
    ((64, 32, 16), (64, 32, 16), 3, 1e-4),
    ((64, 64, 64), (32, 32, 32), 2, 80.0),
    ((64, 64, 64), (64, 64, 64), 1, 3.0),
    ((32, 32, 64), (32, 32, 64), 3, 1e-2),
    ((128, 16, 16), (128, 16, 16), 3, 1e-3),
    ((96, 80, 72), (96, 80, 72), 3, 1e-3),      # tiles interior in x: 128-bit loads of the float volume
    ((50, 44, 40), (50, 44, 40), 3, 1e-3),      # rows of a tile straddle two words of the corrector flags
    ((100, 88, 40), (50, 44, 40), 3, 2e-3),     # the same with four chunks
    # + N(0, sigma) noise (5th field): many outliers, the 1D stream's walker and its one-step
    # single-path descents (speck_dec_fast.cuh, f_walk1d) get real work
    # PWE with magnitudes that need 64 bits: what the forward transform quantised on the fly is
    # discarded, the chunks are transformed again and quantised by k_quantize (pipeline.cu)
    ((32, 32, 32), (32, 32, 32), 3, 1e-10),
    ((64, 32, 32), (32, 32, 32), 3, 2e-10),
    ((64, 64, 64), (64, 64, 64), 3, 1e-3, 3e-4),
    ((64, 64, 64), (32, 32, 32), 3, 1e-3, 1e-3),
]
SYN_GPU = SYN_SMALL + [
    ((256, 256, 128), (256, 256, 128), 3, 1e-3),
    ((512, 64, 64), (512, 64, 64), 2, 90.0),
    ((300, 200, 100), (128, 128, 100), 3, 1e-3),
    ((128, 128, 128), (128, 128, 128), 1, 0.5),
    ((128, 128, 128), (128, 128, 128), 3, 1e-3, 2e-4),
    ((256, 256, 256), (256, 256, 256), 3, 1e-3, 1e-3),   # the noisy bench chunk: 2.6 bpp
]


def syn_id(c):
    return "syn%s-%s-m%d-%g%s" % ("x".join(map(str, c[0])), "x".join(map(str, c[1])), c[2], c[3],
                                  "-noise%g" % c[4] if len(c) > 4 else "")


def syn_field(case):
    v = refs.synthetic_field(case[0], seed=5)
    if len(case) > 4:
        rng = np.random.default_rng(7)
        v = (v + case[4] * rng.standard_normal(v.shape)).astype(v.dtype)
    return v


def check_syn_roundtrip(lib, oracle, case):
    """compress with the library -> bytes equal the oracle's; decompress -> bits equal the oracle's"""
    dims, chunks, mode, q = case[:4]
    v = syn_field(case)
    rc, got = lib.comp_3d(v, dims, chunks, mode, q)
    rc2, exp = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == rc2 == 0
    assert np.array_equal(got, exp)
    check_decomp3d(lib, oracle, exp, True)


def case_id(c):
    return "%s-%s-m%d-%g" % (c[0][:8], "x".join(map(str, c[2])), c[3], c[4])


def fnv1a64(b):
    h = 0xcbf29ce484222325
    for x in bytes(b):
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def check_comp3d(lib, oracle, case, f64=False):
    name, dims, chunks, mode, q = case
    v = refs.load_test_data(name)
    assert v is not None and v.size == dims[0] * dims[1] * dims[2]
    if f64:
        v = v.astype(np.float64)
    rc, got = lib.comp_3d(v, dims, chunks, mode, q)
    rc2, exp = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == rc2 == 0
    assert got.size == exp.size, (got.size, exp.size)
    assert np.array_equal(got, exp), "first diff at byte %d" % int(np.argmax(got != exp))
    return got


def check_decomp3d(lib, oracle, stream, output_float=True):
    rc, got, dims = lib.decomp_3d(stream, output_float)
    rc2, exp, dims2 = oracle.decomp_3d(stream, output_float)
    assert rc == rc2 == 0, (rc, rc2)
    assert dims == dims2
    # bit-identical decoded values
    it = np.uint32 if output_float else np.uint64
    assert np.array_equal(got.view(it), exp.view(it)), "%d values differ" % int(
        np.count_nonzero(got.view(it) != exp.view(it)))
    return got


# ---- 2D slices (sperr_comp_2d / sperr_decomp_2d): (dims, mode, quality, dtype) ----
# fields: a z-plane of refs.synthetic_field (seed 7) cropped to dims, or a golden fixture by name
SLICE_SMALL = [
    ((64, 64), 3, 1e-3, np.float32),
    ((90, 90), 3, 1e-4, np.float32),      # non power of two, 4 transform levels
    ((15, 15), 2, 70.0, np.float32),      # 1 level
    ((8, 8), 3, 1e-2, np.float32),        # no transform: the set I is empty
    ((128, 40), 1, 3.0, np.float32),      # x and y run out of splits at different depths
    ((40, 128), 2, 60.0, np.float64),
    ((33, 70), 1, 20.0, np.float32),      # fixed rate, high-precision retry likely
    ((1, 50), 3, 1e-3, np.float32),       # a single column
    # power-of-two slices take the decoder's fast path (quadtree of aligned boxes + set I)
    ((16, 16), 3, 1e-3, np.float32),
    ((128, 32), 2, 70.0, np.float32),
    ((32, 256), 1, 4.0, np.float32),
    ((256, 256), 3, 1e-4, np.float64),
]
SLICE_GPU = SLICE_SMALL + [
    ((512, 512), 3, 1e-3, np.float32),
    ((1000, 300), 2, 90.0, np.float32),
    ((2048, 2048), 3, 1e-3, np.float32),
    ((256, 256), 1, 1.0, np.float64),
]


def slice_id(c):
    return "slice%dx%d-m%d-%g-%s" % (c[0][0], c[0][1], c[1], c[2], np.dtype(c[3]).name)


def slice_field(dims, dtype, seed=7, nslices=1):
    v = refs.synthetic_field((dims[0], dims[1], nslices), seed=seed, dtype=np.float32)
    return np.ascontiguousarray(v.reshape(nslices, dims[1], dims[0]).astype(dtype))


def check_slice(lib, oracle, case):
    """sperr_comp_2d bytes (with and without header) and sperr_decomp_2d bits equal the oracle's"""
    dims, mode, q, dt = case
    img = slice_field(dims, dt)[0]
    for header in (False, True):
        rc, got = lib.comp_2d(img, dims, mode, q, header)
        rc2, exp = oracle.comp_2d(img, dims, mode, q, header)
        assert rc == rc2 == 0, (rc, rc2)
        assert got.size == exp.size, (got.size, exp.size)
        assert np.array_equal(got, exp), "first diff at byte %d" % int(np.argmax(got != exp))
    rc, exp = oracle.comp_2d(img, dims, mode, q, False)
    for of in (True, False):
        rc, dec = lib.decomp_2d(exp, dims, of)
        rc2, dexp = oracle.decomp_2d(exp, dims, of)
        assert rc == rc2 == 0, (rc, rc2)
        it = np.uint32 if of else np.uint64
        assert np.array_equal(dec.view(it), dexp.view(it)), "%d values differ" % int(
            np.count_nonzero(dec.view(it) != dexp.view(it)))


# ---- multi-resolution decoding: (fixture or synthetic dims, volume dims, chunk dims, mode, quality) ----
MULTIRES_SMALL = [
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 3, 1e-5),   # 3 levels, 2 x 2 x 1 chunks
    ("wmag16.float", (16, 16, 16), (16, 16, 16), 2, 60.0),             # 1 level, one chunk
    ((64, 64, 64), (64, 64, 64), (32, 32, 32), 3, 1e-3),               # synthetic, 2 levels, 8 chunks
    ("wmag17.float", (17, 17, 17), (8, 8, 8), 3, 0.3),                 # not divisible: no hierarchy
]


def check_multires(lib, ref, oracle, case):
    """full volume and every coarse level bit-identical to the reference class (fp64)"""
    name, dims, chunks, mode, q = case
    v = refs.load_test_data(name) if isinstance(name, str) else refs.synthetic_field(name, seed=5)
    rc, stream = oracle.comp_3d(v, dims, chunks, mode, q)
    assert rc == 0
    rc, full, d, levels, ldims = lib.decomp_3d_multires(stream, False)
    if any(a % b for a, b in zip(dims, chunks)):
        # not a whole number of chunks: no hierarchy (the reference's class reads out of bounds here;
        # its CLI refuses the combination, utilities/sperr3d.cpp:236-250), plain decode otherwise
        rc2, exp, d2 = oracle.decomp_3d(stream, False)
        assert rc == rc2 == 0 and d == d2 and levels == []
        assert np.array_equal(full.view(np.uint64), exp.view(np.uint64))
        return 0
    rc2, rfull, rd, rlevels, rldims = ref.decomp_3d_multires(stream)
    assert rc == rc2 == 0 and d == rd
    assert np.array_equal(full.view(np.uint64), rfull.view(np.uint64))
    assert ldims == rldims, (ldims, rldims)
    for a, b in zip(levels, rlevels):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    return len(levels)
