#!/bin/bash
# Round evidence: parity tests, smoke, both bench arms, the 2D slice bench, the launch list of one
# 1024^3 step and one full ncu capture of the dominant kernel of the step (pass its regex as $1,
# default: the per-stream decoder).
K=${1:-k_speck_decode_fast}
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1024.log 2>&1; tail -1 gpurun_out/bench1024.log
timeout 300 python scripts/bench2d.py 256 2048 > gpurun_out/bench2d.log 2>&1; tail -1 gpurun_out/bench2d.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|DeviceRadixSort|DeviceScan|DeviceReduce' -c 1700 --csv --log-file gpurun_out/launches.csv python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_top -f python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
