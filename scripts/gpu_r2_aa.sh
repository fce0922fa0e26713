#!/bin/bash
# round 2, call AA: where the outlier chain is let go (SPERR_B200_GATE none / setup / loop), the whole
# GPU parity suite on the result, the bench as the driver runs it
mkdir -p gpurun_out
Q="python bench.py --steps 6 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', 'step', round(d['ms_per_step'],1), [round(x) for x in d['step_ms_each']], 'c', min(d['compress_ms_each']), 'd', min(d['decompress_ms_each']), ' '.join('%s=%.2f'%(k,s.get(k,-1)) for k in ('c.speck3d','enc.stage_zero','enc.pyramid','enc.lipref_count','enc.plane_loop','enc.lipref_emit','c.inv_quantize','c.idwt','c.outlier_encode')))"; }
for g in none setup loop; do SPERR_B200_GATE=$g $Q 2>/dev/null | show gate_$g; done
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench"
python bench.py > gpurun_out/r2aa_bench1.log 2>&1; tail -1 gpurun_out/r2aa_bench1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d.get('warmup_extra_steps'), d['step_ms_each'], 'c', d['compress_ms_each'], 'd', d['decompress_ms_each'])
print('e2e', d['e2e']); print('stages', d['stages_ms']); print(d['parity'], d['gpu_launches'], d['clocks'])"
