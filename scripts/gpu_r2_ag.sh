#!/bin/bash
# round 2, call AG: sperr_comp_3d from pageable memory with the upload of slab groups overlapped with
# the coder (helper thread through the pinned ring), against upload-then-code; non-temporal stores in
# the staging copy
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decompress.py tests/test_streamed.py -x -q -m gpu 2>&1 | tail -2
run() { env "$@" SPERR_B200_TIMING=1 python bench.py --steps 4 --warmup 2 --settle 0 --cpu-baseline 0 --check 0 2> gpurun_out/r2ag_timing.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$*: pageable %.1f ms (%.2f GB/s)  pinned %.1f ms (%.2f GB/s)' % (e['ms_per_step'], e['value'], e['pinned_source']['ms_per_step'], e['pinned_source']['value']))"
  grep "timing sperr_comp_3d" gpurun_out/r2ag_timing.log | tail -5 | cut -c1-120 | tr '\n' '|'; echo; }
run SPERR_B200_NO_PAGEABLE_OVERLAP=1
run A=1
run SPERR_B200_H2D_NT=1
run SPERR_B200_COPY_THREADS=12
