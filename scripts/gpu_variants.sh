#!/bin/bash
# A/B measurement of library variants in ONE GPU call: variants/<name>.so are copied over the
# product library one after the other, each timed by the same bench line; the last one stays in
# place for a parity run.   usage: scripts/gpu_variants.sh A C B
# Variants are built here, without a GPU:  python -m sperr_b200.build --variant recprefix -DSPERR_REC_PREFIX=1
# (-> variants/recprefix.so; cp sperr_b200/libsperr_b200.so variants/base.so for the baseline) and
# checked for bit-exactness under the emulator first:
#   SPERR_EMUL_DEFS=-DSPERR_REC_PREFIX=1 SPERR_EMUL_TAG=recprefix tests/emul/build.sh
# Restore the product library afterwards (python -m sperr_b200.build --force) and delete variants/.
mkdir -p gpurun_out
for v in "$@"; do
  cp variants/$v.so sperr_b200/libsperr_b200.so
  timeout 120 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 --e2e 0 > gpurun_out/var_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/var_%s.log" % v).read().strip().splitlines()[-1])
    s = d["stages_ms"]
    keys = ["c.stats", "c.dwt", "c.quantize", "c.speck3d", "enc.pyramid", "enc.lipref_count", "enc.plane_loop",
            "enc.lipref_emit", "c.idwt", "c.outlier_encode", "d.speck", "d.reconstruct", "d.outliers", "d.idwt"]
    print(v, "step %.1f ms  comp %.1f  decomp %.1f GB/s |" % (d["ms_per_step"], d["compress_gbs"], d["decompress_gbs"]),
          " ".join("%s=%.2f" % (k, s.get(k, -1)) for k in keys))
except Exception as e:
    print(v, "FAILED", e)
PY
done
timeout 200 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_compress.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
