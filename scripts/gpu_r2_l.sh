#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
export SPERR_B200_DEC_CLUSTER=2
echo "== head trace 128 R=2"; timeout 40 python scripts/dec_trace.py cuda 128 2>&1 | grep "job 0 plane [0-9]*:\|rc" | head -18
for v in cl_unal cl_fence cl_lipl; do echo "== $v"; timeout 40 python scripts/dec_trace.py variants/$v.so 128 2>&1 | grep "rc" | head -3; done
