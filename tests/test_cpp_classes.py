"""include/sperr_b200.hpp (mirrors of sperr::SPERR3D_OMP_C / SPERR3D_OMP_D): compiles with g++
against the C ABI, keeps the reference's return codes (host-only checks, no GPU needed), and -- on a
GPU -- produces the oracle's bytes and bits through the reference's sperr3d flow."""
import os
import subprocess

import numpy as np
import pytest

import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "classes_main")
SO_DIR = os.path.join(ROOT, "sperr_b200")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(os.path.join(SO_DIR, "libsperr_b200.so")):
        from sperr_b200 import build
        build.build()
    src = os.path.join(ROOT, "tests", "cpp", "classes_main.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", SO_DIR, "-lsperr_b200", "-Wl,-rpath," + SO_DIR], check=True)
    return EXE


def test_class_return_codes(exe):
    r = subprocess.run([exe, "hostonly"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout


@pytest.mark.gpu
def test_class_flow_matches_oracle(exe, oracle, tmp_path):
    v = refs.load_test_data("vorticity.128_128_41")
    fin = tmp_path / "in.f32"
    v.astype(np.float32).tofile(fin)
    for mode, q in ((3, 1e-5), (2, 90.0), (1, 3.0)):
        fs, fd = tmp_path / "s.bin", tmp_path / "d.f64"
        r = subprocess.run([exe, "run", str(fin), "128", "128", "41", "64", "64", "41", str(mode), repr(q),
                            str(fs), str(fd)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        rc, exp = oracle.comp_3d(v, (128, 128, 41), (64, 64, 41), mode, q)
        got = np.fromfile(fs, dtype=np.uint8)
        assert np.array_equal(got, exp)
        rc, dexp, dims = oracle.decomp_3d(exp, False)
        assert np.array_equal(np.fromfile(fd, dtype=np.uint64), dexp.view(np.uint64))
