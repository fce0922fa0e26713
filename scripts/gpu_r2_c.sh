#!/bin/bash
# round 2, call C: find the GPU-only decoder failure (emulation is green)
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== R=1, head"; SPERR_B200_DEC_CLUSTER=1 timeout 90 python scripts/dec_debug.py 2>&1 | tail -6
echo "== R=1, old 1D walker"; SPERR_B200_DEC_CLUSTER=1 timeout 90 python scripts/dec_debug.py variants/oldwalk.so 2>&1 | tail -6
echo "== R=2, head"; SPERR_B200_DEC_CLUSTER=2 timeout 90 python scripts/dec_debug.py 2>&1 | tail -6
echo "== racecheck R=1"; SPERR_B200_DEC_CLUSTER=1 timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python scripts/dec_debug.py > gpurun_out/r2c_racecheck.log 2>&1; grep -E "Race|hazard|case|ERROR SUMMARY" gpurun_out/r2c_racecheck.log | head -40
echo "== memcheck R=1"; SPERR_B200_DEC_CLUSTER=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 30 python scripts/dec_debug.py > gpurun_out/r2c_memcheck.log 2>&1; grep -E "Invalid|case|ERROR SUMMARY|at " gpurun_out/r2c_memcheck.log | head -40
