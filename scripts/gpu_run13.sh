#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; tail -1 gpurun_out/bench1024.log | cut -c1-1400
timeout 200 python scripts/decprof.py 256 2>&1 | grep decprof | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv3d -s 4 -c 1 -o gpurun_out/prof_inv3d -f python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_inv.log 2>&1; tail -1 gpurun_out/ncu_inv.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd3d -c 1 -o gpurun_out/prof_fwd3d -f python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_fwd.log 2>&1; tail -1 gpurun_out/ncu_fwd.log | cut -c1-100
