"""Builds sperr_b200/libsperr_b200.so (the product library) in-tree with nvcc for sm_100a.

    python -m sperr_b200.build [--force]
    python -m sperr_b200.build --variant NAME -DFLAG=1 ...   # experiments: variants/NAME.so, built
                                                             # from objects of its own with the
                                                             # extra defines (scripts/gpu_variants.sh)

Every .cu under csrc/ is compiled with ``-gencode arch=compute_100a,code=sm_100a -fmad=false``
(-fmad=false is part of the parity contract: the fp64 lifting / quantiser arithmetic must round
every multiply and add separately, like the reference built with -ffp-contract=off).
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libsperr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("SPERR_B200_PTXAS_V") else "-O3",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _sources():
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))
    return [os.path.join(CSRC, f) for f in srcs]


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


# Per-file ptxas flags (experiments: SPERR_B200_DEC_PTXAS=-O1 compiles the stream decoder at -O1).
PER_FILE = {"speck_dec.cu": os.environ.get("SPERR_B200_DEC_PTXAS")}


def _compile(nvcc, src, obj, extra=()):
    per = PER_FILE.get(os.path.basename(src))
    per = ["-Xptxas", per] if per else []
    cmd = [nvcc] + NVCC_FLAGS + per + list(extra) + os.environ.get("SPERR_B200_EXTRA_NVCC", "").split() + ["-x", "cu", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout[-4000:], r.stderr[-4000:]))
    return r.stderr


def build(force=False, verbose=False, variant=None, extra=()):
    nvcc = _nvcc()
    obj_dir, out = OBJ, OUT
    if variant:
        obj_dir = os.path.join(HERE, "build", "variant_" + variant)
        out = os.path.join(os.path.dirname(HERE), "variants", variant + ".so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    hm = _headers_mtime()
    jobs, objs = [], []
    for s in _sources():
        o = os.path.join(obj_dir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(lambda j: _compile(nvcc, j[0], j[1], extra), jobs):
                if verbose and log:
                    sys.stderr.write(log)
    if jobs or not os.path.exists(out):
        cmd = [nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return out


if __name__ == "__main__":
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose=True, variant=name, extra=defs))
