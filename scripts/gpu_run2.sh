#!/bin/bash
# Second GPU pass: parity tests (compress + decompress), benches, launch list incl. decoder.
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --size 512 --steps 2 --warmup 1 > gpurun_out/bench512.log 2>&1; tail -3 gpurun_out/bench512.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench1024.log 2>&1; tail -3 gpurun_out/bench1024.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches512.csv python bench.py --size 512 --steps 1 --warmup 1 --e2e 0 > gpurun_out/ncu512.log 2>&1; tail -2 gpurun_out/ncu512.log
