#!/bin/bash
# round 2, call V: encoder on a high-priority stream + short-lived outlier-scan CTAs (prio) and the
# reconstruction kernel with per-unit rank bases (recprefix) against the quad-kernel library (new);
# then where the profiler's host stalls come from (ranges by prefix / events without time stamps)
mkdir -p gpurun_out
rm -f variants/base.so
bash scripts/gpu_variants.sh new recprefix prio 2>&1 | grep -v "passed\|^$" | tail -4
echo "== prio with SPERR_B200_NO_PRIO / no segments"
Q="python bench.py --steps 8 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', round(d['ms_per_step'],1), d['step_ms_each'], 'c', min(d['compress_ms_each']), 'd', min(d['decompress_ms_each']), ' '.join('%s=%.2f'%(k,s.get(k,-1)) for k in ('c.speck3d','enc.pyramid','enc.lipref_count','enc.plane_loop','enc.lipref_emit','c.idwt','c.outlier_encode')))"; }
SPERR_B200_NO_PRIO=1 $Q 2>/dev/null | show noprio
SPERR_B200_SCAN_SEG_PAIRS=0 $Q 2>/dev/null | show noseg
SPERR_B200_SCAN_SEG_PAIRS=16 $Q 2>/dev/null | show seg16
echo "== profiler stalls"
for p in c. d. enc dec; do SPERR_B200_PROF_ONLY=$p $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('only $p', d['step_ms_each'])"; done
SPERR_B200_PROF_NOTIMING=1 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('notiming', d['step_ms_each'])"
$Q --diag noprof 2>/dev/null | tail -1
