#!/bin/bash
# 2D parity on the GPU + where the end-to-end time goes (phase timing, PCIe probe, decoder phases)
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_2d.py -x -q > gpurun_out/pytest_gpu_2d.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_2d.log; tail -5 gpurun_out/pytest_gpu_2d.log
timeout 300 python scripts/bench2d.py 64 2048 > gpurun_out/bench2d.log 2>&1; tail -2 gpurun_out/bench2d.log
timeout 300 python scripts/e2e_timing.py 1024 > gpurun_out/e2e_timing.log 2>&1; tail -8 gpurun_out/e2e_timing.log
timeout 200 python scripts/pcie_probe.py > gpurun_out/pcie.log 2>&1; cat gpurun_out/pcie.log
timeout 200 python scripts/decprof.py 256 > gpurun_out/decprof.log 2>&1; tail -6 gpurun_out/decprof.log
