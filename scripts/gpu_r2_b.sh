#!/bin/bash
# round 2, call B: decoder clusters -- parity tests, decoder sweep, bench
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
timeout 1200 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_2d.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
timeout 900 python scripts/dec_sweep.py > gpurun_out/r2b_sweep.log 2>&1; cat gpurun_out/r2b_sweep.log | tail -25
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/r2b_bench1.log 2>&1; tail -1 gpurun_out/r2b_bench1.log | cut -c1-300
