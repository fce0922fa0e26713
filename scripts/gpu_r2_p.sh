#!/bin/bash
# round 2, call P: warp-uniform serial parts + aligned barriers; fall back to the unaligned build if it fails
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
ok=1
for R in 1 2 8; do n=$(SPERR_B200_DEC_CLUSTER=$R timeout 60 python scripts/dec_debug.py 2>&1 | grep -c "rc 0 differing values 0"); echo "small R=$R ok cases: $n"; [ "$n" = 4 ] || ok=0; done
for R in 1 2 8; do r=$(SPERR_B200_DEC_CLUSTER=$R timeout 40 python scripts/dec_trace.py cuda 256 2>&1 | grep "^rc"); echo "256^3 R=$R: $r"; [ "$r" = "rc 0 differing 0" ] || ok=0; done
if [ $ok = 0 ]; then echo "ALIGNED BUILD FAILED -> unaligned variant"; cp variants/unal.so sperr_b200/libsperr_b200.so; fi
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
echo "== sweep"; timeout 300 python scripts/dec_sweep.py 256,512,1024 1,8,auto > gpurun_out/r2p_sweep.log 2>&1; cat gpurun_out/r2p_sweep.log | tail -12
echo "== bench"; timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench1.log 2>&1; tail -1 gpurun_out/r2p_bench1.log | cut -c1-3200
