#!/bin/bash
# round 2, call W: A/B of the L2 prefetch in the inverse transform, the warp-aggregated refinement
# bits of the encoder and the rank-prefix reconstruction; where the host stalls of the step loop sit
mkdir -p gpurun_out
Q="python bench.py --steps 8 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', 'step', round(min(d['step_ms_each']),1), 'c', min(d['compress_ms_each']), 'd', min(d['decompress_ms_each']), ' '.join('%s=%.2f'%(k,s.get(k,-1)) for k in ('c.speck3d','enc.pyramid','enc.lipref_count','enc.plane_loop','enc.lipref_emit','c.idwt','c.outlier_encode','enc1d.lipref_emit','d.speck','d.reconstruct','d.outliers','d.idwt')))
print('   steps', d['step_ms_each']); print('   host ', d['step_host_ms_each']); print('   hostmax', {k:v for k,v in d['stages_host_max_ms'].items() if v > 5})"; }
for v in nopf noagg recprefix head; do
  cp variants/$v.so sperr_b200/libsperr_b200.so
  $Q 2>/dev/null | show $v
done
echo "== parity on head"
timeout 300 python -m pytest tests/test_gpu_compress.py tests/test_gpu_decompress.py tests/test_gpu_2d.py -x -q -m gpu 2>&1 | tail -2
echo "== long step loop"
python bench.py --steps 24 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0 2>/dev/null | show steps24
