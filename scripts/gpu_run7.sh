#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_2d.py tests/test_gpu_decompress.py -x -q > gpurun_out/pytest_gpu_2d.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_2d.log; tail -5 gpurun_out/pytest_gpu_2d.log
timeout 300 python scripts/bench2d.py 64 2048 > gpurun_out/bench2d.log 2>&1; tail -2 gpurun_out/bench2d.log
timeout 300 python scripts/bench2d.py 256 2048 > gpurun_out/bench2d_256.log 2>&1; tail -1 gpurun_out/bench2d_256.log
