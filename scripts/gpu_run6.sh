#!/bin/bash
mkdir -p gpurun_out
for v in plain keep nccl devfirst nccl_devfirst_keep; do
timeout 200 python scripts/e2e_probe.py $v 2>&1 | grep -v "^\[" | tail -7
done > gpurun_out/e2e_probe.log 2>&1
cat gpurun_out/e2e_probe.log
