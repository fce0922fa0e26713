#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_2d.py -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_d.log; tail -3 gpurun_out/pytest_gpu_d.log
timeout 300 python scripts/e2e_timing.py 1024 2>&1 | grep -v "^\[" | tail -4
SPERR_B200_NO_NT_COPY=1 timeout 300 python scripts/e2e_timing.py 1024 2>&1 | grep -v "^\[" | tail -2
SPERR_B200_COPY_THREADS=8 timeout 300 python scripts/e2e_timing.py 1024 2>&1 | grep -v "^\[" | tail -2
