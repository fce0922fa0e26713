#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== head"; timeout 60 python scripts/dec_trace.py cuda 2>&1 | grep "lists\|walk:\|rc" | head -12
echo "== zero smem"; timeout 60 python scripts/dec_trace.py variants/zerosmem.so 2>&1 | grep "lists\|walk:\|rc" | head -12
echo "== round1 lib"; SPERR_B200_DEC_CLUSTER=1 timeout 90 python scripts/dec_debug.py variants/round1.so 2>&1 | tail -5
