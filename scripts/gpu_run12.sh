#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_compress.py -x -q > gpurun_out/pytest_gpu_comp.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_comp.log; tail -3 gpurun_out/pytest_gpu_comp.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; tail -1 gpurun_out/bench1024.log | cut -c1-1600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lipref_emit -s 2 -c 1 -o gpurun_out/prof_emit -f python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_emit.log 2>&1; tail -1 gpurun_out/ncu_emit.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 100 -c 1 -o gpurun_out/prof_expand -f python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_expand.log 2>&1; tail -1 gpurun_out/ncu_expand.log | cut -c1-100
