#!/bin/bash
# round 2, call Z (2 GPUs): the multi-device C API after the per-device read-back fix
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; tail -15 gpurun_out/r2z_pytest.log | cut -c1-300
timeout 600 python scripts/e2e_multi.py 1024 2>&1 | tail -6 | cut -c1-300
