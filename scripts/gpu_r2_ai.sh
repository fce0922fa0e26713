#!/bin/bash
# round 2, call AI: forward transform with shared-space accesses and 32-bit indices (parity + c.dwt);
# does the clock sampler cost steps? (default against --diag noclocks, 10 steps each, twice)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compress.py tests/test_gpu_fullsize.py tests/test_fma_flavour.py -x -q -m gpu 2>&1 | tail -2
Q="python bench.py --steps 10 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0"
for i in 1 2; do
  $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default ', round(d['ms_per_step'],1), [round(x,1) for x in d['step_ms_each']], 'c.dwt', d['stages_ms']['c.dwt'], d['host_counters_delta'], d['clocks']['samples'])"
  $Q --diag noclocks 2>/dev/null | tail -1 | cut -c1-400
done
