#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
for lib in cuda variants/unal.so; do
for R in 1 2 8; do n=$(SPERR_B200_DEC_CLUSTER=$R timeout 60 python scripts/dec_debug.py $lib 2>&1 | grep -c "rc 0 differing values 0"); echo "$lib small R=$R ok cases: $n"; done
for R in 1 2 8; do r=$(SPERR_B200_DEC_CLUSTER=$R timeout 40 python scripts/dec_trace.py $lib 256 2>&1 | grep "^rc"); echo "$lib 256^3 R=$R: $r"; done
done
