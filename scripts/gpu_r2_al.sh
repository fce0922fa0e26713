#!/bin/bash
# round 2, call AL (last GPU minutes): PWE with quantisation fused into the forward transform and
# de-quantisation into the outlier scan -- parity suite, then A/B of one bench line each way
mkdir -p gpurun_out
echo "== pytest"; timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2al_pytest.log 2>&1; tail -3 gpurun_out/r2al_pytest.log | cut -c1-300
Q="python bench.py --steps 5 --warmup 3 --settle 0.5 --e2e 0 --cpu-baseline 0"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', 'step', round(min(d['step_ms_each']),1), [round(x) for x in d['step_ms_each']], 'c', min(d['compress_ms_each']), 'd', min(d['decompress_ms_each']), d.get('parity',{}).get('matches_recorded_1gpu_fingerprint'))
print('   ', ' '.join('%s=%.2f'%(k,v) for k,v in sorted(s.items()) if k.startswith('c.') or k.startswith('enc.')))"; }
SPERR_B200_NO_FUSED_QUANT=1 SPERR_B200_NO_FUSED_DEQ=1 $Q 2>/dev/null | show separate
$Q 2>/dev/null | show fused
SPERR_B200_KEEP_COEF=1 $Q 2>/dev/null | show fused_keepcoef
