#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
for lib in cuda variants/o3u.so; do
for R in 1 2 8; do echo "== $lib small R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 60 python scripts/dec_debug.py $lib 2>&1 | grep -c "rc 0 differing values 0"; done
for R in 2 8; do echo "== $lib 128 R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 40 python scripts/dec_trace.py $lib 128 2>&1 | grep "^rc"; done
for R in 1 2 8; do echo "== $lib 256 R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 40 python scripts/dec_trace.py $lib 256 2>&1 | grep "^rc"; done
done
