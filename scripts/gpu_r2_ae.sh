#!/bin/bash
# round 2, call AE (2 GPUs): final state. Whole GPU parity suite (with the two-GPU tests and the
# streamed path), smoke, the bench as the driver runs it at N=1 and N=2, the multi-device C API
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_pytest.log 2>&1; tail -4 gpurun_out/r2ae_pytest.log | cut -c1-300
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
unset SPERR_B200_VERBOSE
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'steps', d['step_ms_each'], 'c', d['compress_ms_each'], 'd', d['decompress_ms_each'], 'extra warm-up', d.get('warmup_extra_steps'))
print('   ', ' '.join('%s=%.2f'%(k,v) for k,v in sorted(s.items())))
print('    e2e', d.get('e2e'), 'parity', d.get('parity'), 'launches', d.get('gpu_launches'))
print('    roofline', d.get('roofline')); print('    wavelet', d.get('roofline_wavelet')); print('    cpu', d.get('cpu_baseline'))
print('    extra', json.dumps(d.get('extra'))[:500])"; }
echo "== bench N=1"
python bench.py > gpurun_out/r2ae_bench1.log 2>&1; tail -1 gpurun_out/r2ae_bench1.log | show n1
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2ae_bench2.log 2>&1; tail -1 gpurun_out/r2ae_bench2.log | show n2
echo "== multi-device C API"
timeout 600 python scripts/e2e_multi.py 1024 2>&1 | tail -3 | cut -c1-300
