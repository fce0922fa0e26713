#!/bin/bash
# round 2, call AB: host copy threads of the host-pointer API (pageable source) -- fewer threads leave
# cores to the coder's launch threads
mkdir -p gpurun_out
nproc
for t in 14 10 8 6 4; do
  SPERR_B200_COPY_THREADS=$t python bench.py --steps 3 --warmup 1 --settle 0 --cpu-baseline 0 --check 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('copy threads $t: pageable %.1f ms (%.2f GB/s)  pinned %.1f ms (%.2f GB/s)' % (e['ms_per_step'], e['value'], e['pinned_source']['ms_per_step'], e['pinned_source']['value']))"
done
SPERR_B200_TIMING=1 python bench.py --steps 2 --warmup 1 --settle 0 --cpu-baseline 0 --check 0 2>&1 | grep "sperr_b200 timing" | tail -8 | cut -c1-300
