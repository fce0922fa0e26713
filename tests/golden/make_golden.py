#!/usr/bin/env python
"""Generates tests/golden/reference_fingerprints.json with the UNMODIFIED reference.

Run in the build container (needs /root/reference; oracle/Makefile compiles it, STRICT flavour =
-ffp-contract=off, into oracle/_ref/libsperr_ref.so):

    python tests/golden/make_golden.py

For every case below the reference's own C API (sperr_comp_3d / sperr_decomp_3d / sperr_comp_2d /
sperr_decomp_2d, /root/reference/src/SPERR_C_API.cpp) compresses a committed fixture and decodes
the result; the JSON keeps the stream length and the FNV-1a 64 hash of the stream bytes and of the
decoded values. tests/test_golden_kats.py holds the oracle (CPU) and the CUDA library (GPU) to
these records, so the oracle stays pinned where the reference cannot travel (the GPU box).
Inputs are files only (no generated fields: their bits would depend on the local libm).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import refs   # noqa: E402

# (fixture, dims, chunk dims, mode, quality, input dtype)
CASES_3D = [c + ("f32",) for c in cases.COMP3D_GPU] + [
    # the stream fingerprints SURVEY.md 8c recorded from the reference during the survey
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 2, 125.0, "f32"),
    # double input (the reference converts nothing: fp64 in, fp64 arithmetic)
    ("wmag17.float", (17, 17, 17), (17, 17, 17), 3, 0.3, "f64"),
    ("vorticity.128_128_41", (128, 128, 41), (64, 64, 41), 2, 100.0, "f64"),
]
# (fixture, dims, mode, quality)
CASES_2D = [
    ("15x15.float", (15, 15), 2, 70.0),
    ("15x15.float", (15, 15), 3, 1e-2),
    ("90x90.float", (90, 90), 3, 1e-4),
    ("90x90.float", (90, 90), 2, 80.0),
    ("90x90.float", (90, 90), 1, 3.0),
]


def key3d(c):
    return "3d|%s|%s|%s|m%d|%r|%s" % (c[0], "x".join(map(str, c[1])), "x".join(map(str, c[2])), c[3], c[4], c[5])


def key2d(c):
    return "2d|%s|%s|m%d|%r" % (c[0], "x".join(map(str, c[1])), c[2], c[3])


def load3d(c):
    v = refs.load_test_data(c[0])
    return v.astype(np.float64) if c[5] == "f64" else v


def main():
    ref = refs.ref()
    assert ref is not None, "oracle/_ref/libsperr_ref.so not built (needs /root/reference)"
    out = {}
    for c in CASES_3D:
        v = load3d(c)
        rc, s = ref.comp_3d(v, c[1], c[2], c[3], c[4])
        assert rc == 0, (c, rc)
        rc, d32, dims = ref.decomp_3d(s, True)
        assert rc == 0
        rc, d64, _ = ref.decomp_3d(s, False)
        assert rc == 0
        out[key3d(c)] = {"len": int(s.size), "stream": cases.fnv1a64(s), "dec_f32": cases.fnv1a64(d32.tobytes()),
                         "dec_f64": cases.fnv1a64(d64.tobytes())}
    for c in CASES_2D:
        img = refs.load_test_data(c[0])
        rc, s = ref.comp_2d(img, c[1], c[2], c[3], header=False)
        assert rc == 0, (c, rc)
        rc, d32 = ref.decomp_2d(s, c[1], True)[:2]
        assert rc == 0
        out[key2d(c)] = {"len": int(s.size), "stream": cases.fnv1a64(s), "dec_f32": cases.fnv1a64(d32.tobytes())}
    with open(os.path.join(HERE, "reference_fingerprints.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "NCAR/SPERR 0.8.5, STRICT (-ffp-contract=off)",
                   "hash": "FNV-1a 64", "records": out}, f, indent=1, sort_keys=True)
    print("wrote %d records" % len(out))


if __name__ == "__main__":
    main()
