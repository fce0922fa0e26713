#!/bin/bash
# First GPU pass: parity tests, a short bench, kernel launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt; free -g >> gpurun_out/lscpu.txt
export SPERR_B200_VERBOSE=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --size 512 --steps 2 --warmup 1 > gpurun_out/bench512.log 2>&1; tail -3 gpurun_out/bench512.log
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/bench1024.log 2>&1; tail -3 gpurun_out/bench1024.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches512.csv python bench.py --size 512 --steps 1 --warmup 1 --e2e 0 > gpurun_out/ncu512.log 2>&1; tail -2 gpurun_out/ncu512.log
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.log 2>&1; tail -2 gpurun_out/bench_ref.log
