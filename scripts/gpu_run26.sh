#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024_$i.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench1024_$i.log').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],1), round(d['compress_gbs'],1), round(d['decompress_gbs'],1), d['e2e']['value'], d['clocks'])
PY
done
