#!/bin/bash
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
ok=1
for R in 1 2 8; do echo "== O1 dec R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 90 python scripts/dec_debug.py 2>&1 | tail -5 | tee gpurun_out/r2j_dec$R.log; grep -q "rc -1\|differing values [1-9]" gpurun_out/r2j_dec$R.log && ok=0; done
for R in 1 8; do echo "== O2 dec R=$R"; SPERR_B200_DEC_CLUSTER=$R timeout 90 python scripts/dec_debug.py variants/dec_o2.so 2>&1 | tail -5; done
if [ $ok = 0 ]; then echo "decoder still broken: stopping"; exit 0; fi
echo "== pytest"; timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log
echo "== sweep"; timeout 400 python scripts/dec_sweep.py > gpurun_out/r2j_sweep.log 2>&1; grep -v decprof gpurun_out/r2j_sweep.log | tail -20
echo "== bench"; timeout 400 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/r2j_bench1.log 2>&1; tail -1 gpurun_out/r2j_bench1.log | cut -c1-2500
echo "== bench no TMA"; SPERR_B200_NO_TMA=1 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 --e2e 0 --check 0 > gpurun_out/r2j_bench_notma.log 2>&1; tail -1 gpurun_out/r2j_bench_notma.log | cut -c1-1200
