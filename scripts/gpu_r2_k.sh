#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
for R in 1 2 4 8; do echo "== 128^3 R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 100 python scripts/dec_big.py 128 2>&1 | tail -2; done
for R in 2 8; do echo "== 256^3 R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 100 python scripts/dec_big.py 256 2>&1 | tail -2; done
echo "== memcheck 128^3 R=2"; SPERR_B200_DEC_CLUSTER=2 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dec_big.py 128 > gpurun_out/r2k_memcheck.log 2>&1; grep -E "Invalid|Out-of|at sperr|by thread|dims|ERROR SUMMARY" gpurun_out/r2k_memcheck.log | head -20
echo "== racecheck 128^3 R=2"; SPERR_B200_DEC_CLUSTER=2 timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 50 python scripts/dec_big.py 128 > gpurun_out/r2k_racecheck.log 2>&1; grep -E "Thread|dims|SUMMARY" gpurun_out/r2k_racecheck.log | sed 's/.*at sperr_b200:://' | sort | uniq -c | sort -rn | head -20
