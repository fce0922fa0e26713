"""tools/sperr3d and tools/sperr2d (front ends flag-compatible with the reference's utilities,
/root/reference/utilities/sperr3d.cpp, sperr2d.cpp): argument checks without a GPU; on a GPU the
files they write are the oracle's bytes / bits (BASELINE.json config #1: "sperr3d CLI
compress+decompress ... PWE tolerance, single chunk")."""
import os
import re
import subprocess

import numpy as np
import pytest

import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO_DIR = os.path.join(ROOT, "sperr_b200")


def _build(name):
    if not os.path.exists(os.path.join(SO_DIR, "libsperr_b200.so")):
        from sperr_b200 import build
        build.build()
    exe = os.path.join(ROOT, "tools", name)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tools", name + ".cpp"), "-o", exe, "-L", SO_DIR, "-lsperr_b200",
                    "-Wl,-rpath," + SO_DIR], check=True)
    return exe


@pytest.fixture(scope="module")
def sperr3d():
    return _build("sperr3d")


@pytest.fixture(scope="module")
def sperr2d():
    return _build("sperr2d")


def run(*a):
    r = subprocess.run(list(a), capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def test_sperr3d_argument_checks(sperr3d):
    f = os.path.join(ROOT, "tests", "golden", "wmag17.float")
    assert "What's the input file?" in run(sperr3d)[1]
    assert "compressing (-c) or decompressing (-d)" in run(sperr3d, f)[1]
    assert "dimensions of this 3D volume" in run(sperr3d, "-c", f)[1]
    assert "floating-type precision" in run(sperr3d, "-c", "--dims", "17", "17", "17", f)[1]
    assert "compression quality" in run(sperr3d, "-c", "--dims", "17", "17", "17", "--ftype", "32", f)[1]
    rc, out = run(sperr3d, "-c", "--dims", "17", "17", "16", "--ftype", "32", "--pwe", "0.1", f)
    assert rc != 0 and "Input file size wrong!" in out
    assert "needs an output destination" in run(sperr3d, "-d", f)[1]
    assert run(sperr3d, "-c", "-d", f)[0] != 0
    assert run(sperr3d, "-c", "--pwe", "1", "--psnr", "2", f)[0] != 0
    assert run(sperr3d, "--help")[0] == 0


def test_sperr2d_argument_checks(sperr2d):
    f = os.path.join(ROOT, "tests", "golden", "15x15.float")
    assert "dimensions of this 2D slice" in run(sperr2d, "-c", f)[1]
    rc, out = run(sperr2d, "-c", "--dims", "15", "14", "--ftype", "32", "--pwe", "0.1", f)
    assert rc != 0 and "Input file size wrong!" in out


@pytest.mark.gpu
def test_sperr3d_files_match_oracle(sperr3d, oracle, tmp_path):
    for name, dims, chunks, flag, q, mode in (("wmag17.float", (17, 17, 17), (17, 17, 17), "--pwe", 0.3, 3),
                                              ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), "--psnr", 90.0, 2),
                                              ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), "--bpp", 3.0, 1)):
        src = os.path.join(ROOT, "tests", "golden", name)
        bs, df, dd, df2 = (str(tmp_path / n) for n in ("s.sperr", "d.f32", "d.f64", "d2.f32"))
        rc, out = run(sperr3d, "-c", "--ftype", "32", "--dims", *map(str, dims), "--chunks", *map(str, chunks),
                      flag, repr(q), "--bitstream", bs, "--decomp_d", dd, "--print_stats", src)
        assert rc == 0, out
        assert re.search(r"Input range = \(.*\), L-Infty = ", out) and "Bitrate = " in out and "PSNR = " in out
        v = refs.load_test_data(name)
        rc2, exp = oracle.comp_3d(v, dims, chunks, mode, q)
        assert rc2 == 0 and np.array_equal(np.fromfile(bs, dtype=np.uint8), exp)
        rc, out = run(sperr3d, "-d", "--decomp_f", df, "--decomp_d", dd, bs)
        assert rc == 0, out
        _, dexp, _ = oracle.decomp_3d(exp, True)
        _, dexp64, _ = oracle.decomp_3d(exp, False)
        assert np.array_equal(np.fromfile(df, dtype=np.uint32), dexp.view(np.uint32))
        assert np.array_equal(np.fromfile(dd, dtype=np.uint64), dexp64.view(np.uint64))
        if mode == 3:   # the printed L-infinity respects the tolerance
            linf = float(re.search(r"L-Infty = ([0-9.eE+-]+)", run(
                sperr3d, "-c", "--ftype", "32", "--dims", *map(str, dims), "--chunks", *map(str, chunks), flag,
                repr(q), "--print_stats", src)[1]).group(1))
            assert linf <= q * 1.01


@pytest.mark.gpu
def test_sperr2d_files_match_oracle(sperr2d, oracle, tmp_path):
    src = os.path.join(ROOT, "tests", "golden", "90x90.float")
    v = np.fromfile(src, dtype=np.float32)
    bs, df = str(tmp_path / "s.sperr"), str(tmp_path / "d.f32")
    rc, out = run(sperr2d, "-c", "--ftype", "32", "--dims", "90", "90", "--pwe", "0.01", "--bitstream", bs, src)
    assert rc == 0, out
    rc2, exp = oracle.comp_2d(v, (90, 90), 3, 0.01, True)
    assert rc2 == 0 and np.array_equal(np.fromfile(bs, dtype=np.uint8), exp)
    rc, out = run(sperr2d, "-d", "--decomp_f", df, bs)
    assert rc == 0, out
    _, dexp = oracle.decomp_2d(exp[10:], (90, 90), True)
    assert np.array_equal(np.fromfile(df, dtype=np.uint32), dexp.view(np.uint32))


@pytest.mark.gpu
def test_sperr3d_lowres_files(sperr3d, oracle, tmp_path):
    """--decomp_lowres_d: one file per coarsened level, named like the reference's, holding what the
    reference class returns (oracle/ref_shim.cpp)"""
    ref = refs.ref()
    if ref is None or not hasattr(ref.lib, "ref_decomp_3d_multires"):
        pytest.skip("reference library with the multi-resolution shim not built")
    src = os.path.join(ROOT, "tests", "golden", "vorticity.128_128_41")
    bs, low = str(tmp_path / "s.sperr"), str(tmp_path / "low")
    rc, out = run(sperr3d, "-c", "--ftype", "32", "--dims", "128", "128", "41", "--chunks", "64", "64", "41",
                  "--pwe", "1e-5", "--bitstream", bs, src)
    assert rc == 0, out
    rc, out = run(sperr3d, "-d", "--decomp_lowres_d", low, bs)
    assert rc == 0, out
    rc2, full, d, levels, ldims = ref.decomp_3d_multires(np.fromfile(bs, dtype=np.uint8))
    assert rc2 == 0 and len(levels) == 3
    for lv, ld in zip(levels, ldims):
        f = "%s.%dx%dx%d" % (low, ld[0], ld[1], ld[2])
        assert np.array_equal(np.fromfile(f, dtype=np.uint64), lv.view(np.uint64)), f
    rc, out = run(sperr3d, "-c", "--ftype", "32", "--dims", "128", "128", "41", "--chunks", "60", "64", "41",
                  "--pwe", "1e-5", "--decomp_lowres_f", low, src)
    assert rc != 0 and "cannot support multi-resolution decoding" in out
