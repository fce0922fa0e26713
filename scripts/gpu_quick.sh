#!/bin/bash
# quick check after a change: sharded + decoder parity tests, one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench1024.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['compress_gbs'], d['decompress_gbs'], d['e2e']['value'])
PY
done
