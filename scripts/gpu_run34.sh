#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress.py tests/test_gpu_decompress.py -x -q 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k 'regex:k_fwd3d' -c 5 --csv --log-file gpurun_out/l_fwd.csv python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > /dev/null 2>&1
grep -E "k_fwd3d" gpurun_out/l_fwd.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-90 | head -10
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench1024.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['compress_gbs'], d['decompress_gbs'], d['e2e']['value'], d['stages_ms'])
PY
