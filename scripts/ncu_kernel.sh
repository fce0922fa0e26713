#!/bin/bash
# usage: ncu_kernel.sh <kernel regex> <out name> [skip] [count]  -- full ncu capture on the 512^3 bench
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-0} -c ${4:-1} -o gpurun_out/$2 -f python bench.py --size 512 --steps 1 --warmup 0 --e2e 0 --cpu-baseline 0 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log | cut -c1-200
