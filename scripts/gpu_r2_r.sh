#!/bin/bash
# round 2, call R: full parity suite, decoder sweep, bench, e2e batching sweep
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; tail -4 gpurun_out/r2r_pytest.log
echo "== sweep"; timeout 300 python scripts/dec_sweep.py 256,512,1024 1,2,8,auto > gpurun_out/r2r_sweep.log 2>&1; grep -v decprof gpurun_out/r2r_sweep.log | tail -14; grep decprof gpurun_out/r2r_sweep.log | head -2
echo "== bench"; timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench1.log 2>&1; tail -1 gpurun_out/r2r_bench1.log | cut -c1-3400
echo "== bench stagger"; SPERR_B200_STAGGER=1 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 --e2e 0 --check 0 > gpurun_out/r2r_bench_stagger.log 2>&1; tail -1 gpurun_out/r2r_bench_stagger.log | cut -c1-1300
for mc in 32 16; do echo "== e2e overlap min chunks $mc"; SPERR_B200_OVERLAP_MIN_CHUNKS=$mc timeout 200 python scripts/e2e_timing.py 2>&1 | tail -2; done
echo "== e2e default"; timeout 200 python scripts/e2e_timing.py 2>&1 | tail -2
