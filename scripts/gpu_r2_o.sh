#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== failing case"; timeout 60 python scripts/one_case.py 2>&1 | tail -4
echo "== failing case, no TMA"; SPERR_B200_NO_TMA=1 timeout 60 python scripts/one_case.py 2>&1 | tail -4
echo "== memcheck"; timeout 200 compute-sanitizer --tool memcheck --print-limit 4 python scripts/one_case.py > gpurun_out/r2o_memcheck.log 2>&1; grep -E "Invalid|Out-of|Illegal|at sperr|at void sperr|by thread|rc|ERROR SUMMARY" gpurun_out/r2o_memcheck.log | head -16
echo "== O1 unaligned decode timing"; cp variants/o1u.so sperr_b200/libsperr_b200.so; timeout 200 python scripts/dec_sweep.py 256,1024 1,8 2>&1 | grep -v decprof | tail -4
