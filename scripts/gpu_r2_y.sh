#!/bin/bash
# round 2, call Y (2 GPUs): staging array cleared by dirty ranges (1-GPU bench + compress parity),
# one call over two GPUs behind the C API (SPERR_B200_DEVICES), NCCL world-2 byte identity,
# strong-scaled bench at N=2 with the reference arm
mkdir -p gpurun_out
nvidia-smi -L | head -3
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$1', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'steps', d['step_ms_each'], 'c', min(d['compress_ms_each']), 'd', min(d['decompress_ms_each']))
print('   ', ' '.join('%s=%.2f'%(k,v) for k,v in sorted(s.items())))
print('    e2e', d.get('e2e'), 'parity', d.get('parity'), 'extra', json.dumps(d.get('extra'))[:600])"; }
echo "== 1 GPU"
python bench.py --steps 5 --warmup 3 --e2e 0 --cpu-baseline 0 2>/dev/null | show n1
echo "== parity (1 GPU tests + multi-device tests)"
timeout 900 python -m pytest tests/test_gpu_compress.py tests/test_gpu_2d.py tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
echo "== 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2y_bench2.log 2>&1; tail -1 gpurun_out/r2y_bench2.log | show n2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2y_ref2.log 2>&1; tail -1 gpurun_out/r2y_ref2.log | cut -c1-300
