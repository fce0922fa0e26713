#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
for v in vol noinl finl o1; do echo "== $v"; timeout 60 python scripts/dec_trace.py variants/$v.so 2>&1 | grep "walk: plane 31\|rc" | head -4; done
