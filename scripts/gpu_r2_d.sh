#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
export DEC_DEBUG_ONE=1
echo "== synccheck R=1"; SPERR_B200_DEC_CLUSTER=1 timeout 300 compute-sanitizer --tool synccheck --print-limit 40 python scripts/dec_debug.py > gpurun_out/r2d_synccheck.log 2>&1; grep -E "Barrier|barrier|at |case|ERROR SUMMARY" gpurun_out/r2d_synccheck.log | head -40
echo "== racecheck R=1 (more)"; SPERR_B200_DEC_CLUSTER=1 timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 400 python scripts/dec_debug.py > gpurun_out/r2d_racecheck.log 2>&1; grep -E "Thread|case|SUMMARY" gpurun_out/r2d_racecheck.log | sed 's/.*at sperr_b200:://' | sort | uniq -c | sort -rn | head -30
