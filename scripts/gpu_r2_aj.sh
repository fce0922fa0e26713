#!/bin/bash
# round 2, call AJ: the bench with the clocks read by the timing thread (no sampler thread): ten steps,
# then the default run as the driver makes it
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ten steps', round(d['ms_per_step'],1), [round(x,1) for x in d['step_ms_each']], d['clocks'], d['host_counters_delta'])"
python bench.py > gpurun_out/r2aj_bench1.log 2>&1; tail -1 gpurun_out/r2aj_bench1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d.get('warmup_extra_steps'), d['step_ms_each'], 'c', d['compress_ms_each'], 'd', d['decompress_ms_each'])
print('e2e', d['e2e']); print('stages', d['stages_ms']); print(d['parity'], d['gpu_launches'], d['clocks']); print(d['roofline_wavelet'])"
