#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 300 python scripts/bench2d.py 64 2048 > gpurun_out/bench2d.log 2>&1; tail -2 gpurun_out/bench2d.log
SPERR_B200_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_timing.log 2>&1; grep -c timing gpurun_out/bench_timing.log; grep timing gpurun_out/bench_timing.log | tail -8; tail -1 gpurun_out/bench_timing.log | cut -c1-400
