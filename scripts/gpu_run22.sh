#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/e2e_timing.py 1024 2>&1 | grep -v "^\[" | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; tail -1 gpurun_out/bench1024.log | cut -c1-1400
