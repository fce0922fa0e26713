#!/bin/bash
# round 2, call AK: final code -- GPU parity suite, smoke, the bench as the driver runs it
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ak_pytest.log 2>&1; tail -3 gpurun_out/r2ak_pytest.log | cut -c1-300
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench"
python bench.py > gpurun_out/r2ak_bench1.log 2>&1; tail -1 gpurun_out/r2ak_bench1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d.get('warmup_extra_steps'), d['step_ms_each'], 'c', d['compress_ms_each'], 'd', d['decompress_ms_each'])
print('host', d['step_host_ms_each'], d['host_counters_delta'])
print('e2e', d['e2e']); print('stages', d['stages_ms']); print(d['parity'], d['gpu_launches'], d['clocks']); print(d['cpu_baseline'])"
