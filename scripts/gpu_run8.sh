#!/bin/bash
mkdir -p gpurun_out
for m in 16 32; do
echo "== SPERR_B200_OVERLAP_MIN_CHUNKS=$m"
SPERR_B200_OVERLAP_MIN_CHUNKS=$m timeout 300 python scripts/e2e_timing.py 1024 2>&1 | grep -v "^\[" | tail -6
done > gpurun_out/overlap_probe.log 2>&1
cat gpurun_out/overlap_probe.log
