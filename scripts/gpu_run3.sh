#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench1024.log 2>&1; tail -3 gpurun_out/bench1024.log
