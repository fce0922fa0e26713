#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== dec trace R=1"; timeout 60 python scripts/dec_trace.py > gpurun_out/r2f_trace.log 2>&1; head -24 gpurun_out/r2f_trace.log
echo "== enc"; timeout 60 python scripts/enc_debug.py 2>&1 | tail -6
echo "== enc no TMA"; SPERR_B200_NO_TMA=1 timeout 60 python scripts/enc_debug.py 2>&1 | tail -6
echo "== enc memcheck"; ENC_DEBUG_ONE=2 timeout 200 compute-sanitizer --tool memcheck --print-limit 10 python scripts/enc_debug.py > gpurun_out/r2f_memcheck.log 2>&1; grep -E "Invalid|Out-of|at |by thread|case|ERROR SUMMARY" gpurun_out/r2f_memcheck.log | head -30
