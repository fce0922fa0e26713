#!/bin/bash
# round 2, call AC: the 1D walker's one-step single-path descents (walkfast) against the bit-by-bit
# walker (nowalkfast): decoder stage time at 1 / 8 / 64 chunks, noisy field, GPU parity suite
mkdir -p gpurun_out
for v in nowalkfast walkfast; do
  cp variants/$v.so sperr_b200/libsperr_b200.so
  echo "== $v"
  timeout 600 python scripts/dec_sweep.py 256,512,1024 auto 2>&1 | grep -E "^n=|decprof job 1" | cut -c1-200
  timeout 300 python scripts/dec_noisy.py 512 2>&1 | grep -E "^R=auto|^n=" | cut -c1-300
done
echo "== pytest (walkfast)"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1; tail -3 gpurun_out/r2ac_pytest.log
