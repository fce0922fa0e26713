#!/bin/bash
# round 2, call AD: LIP passes that skip the 128-word groups of the mask nothing was ever put in
# (lipsum) against full sweeps (nolipsum): decoder stage and phase cycles, parity
mkdir -p gpurun_out
for v in nolipsum lipsum; do
  cp variants/$v.so sperr_b200/libsperr_b200.so
  echo "== $v"
  timeout 600 python scripts/dec_sweep.py 256,512,1024 1,auto 2>&1 | grep -E "^n=|decprof job 0" | cut -c1-200
done
echo "== pytest (lipsum)"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_pytest.log 2>&1; tail -3 gpurun_out/r2ad_pytest.log
