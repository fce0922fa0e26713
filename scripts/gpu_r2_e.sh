#!/bin/bash
# round 2, call E: decoder after the L2-load fix; encoder persistent plane kernel; TMA forward transform
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== dec R=1"; SPERR_B200_DEC_CLUSTER=1 timeout 120 python scripts/dec_debug.py 2>&1 | tail -5
echo "== dec R=2"; SPERR_B200_DEC_CLUSTER=2 timeout 120 python scripts/dec_debug.py 2>&1 | tail -5
echo "== dec R=8"; SPERR_B200_DEC_CLUSTER=8 timeout 120 python scripts/dec_debug.py 2>&1 | tail -5
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
echo "== sweep"; timeout 600 python scripts/dec_sweep.py > gpurun_out/r2e_sweep.log 2>&1; grep -v decprof gpurun_out/r2e_sweep.log | tail -20
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/r2e_bench1.log 2>&1; tail -1 gpurun_out/r2e_bench1.log | cut -c1-2500
echo "== bench no TMA"; SPERR_B200_NO_TMA=1 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 --e2e 0 --check 0 > gpurun_out/r2e_bench_notma.log 2>&1; tail -1 gpurun_out/r2e_bench_notma.log | cut -c1-1500
