#!/bin/bash
# round 2, call AF: group size of the batched upload in sperr_comp_3d (chunks per group) with the
# faster coder, pageable and pinned source
mkdir -p gpurun_out
for g in 16 32 64; do
  SPERR_B200_OVERLAP_MIN_CHUNKS=$g SPERR_B200_TIMING=1 python bench.py --steps 4 --warmup 2 --settle 0 --cpu-baseline 0 --check 0 2> gpurun_out/r2af_timing_$g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('min group $g: pageable %.1f ms (%.2f GB/s)  pinned %.1f ms (%.2f GB/s)' % (e['ms_per_step'], e['value'], e['pinned_source']['ms_per_step'], e['pinned_source']['value']))"
  grep "timing sperr_comp_3d" gpurun_out/r2af_timing_$g.log | tail -6 | cut -c1-200
done
