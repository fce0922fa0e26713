#!/bin/bash
# round 2, call U: inverse transform with one quad per thread against the previous kernel (A/B), the
# whole GPU parity suite on the new library, step times with the pooled profiler events, ncu of the
# new kernel, noisy config 2
mkdir -p gpurun_out /tmp/rep
bash scripts/gpu_variants.sh base new 2>&1 | tail -4
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
echo "== bench"
python bench.py --steps 8 --warmup 3 > gpurun_out/r2u_bench1.log 2>&1; tail -1 gpurun_out/r2u_bench1.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['step_ms_each'], d['compress_ms_each'], d['decompress_ms_each'], d['e2e'], d['stages_ms'])"
B="python bench.py --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 --check 0"
cap() {  # name, regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -c $cnt -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1
  python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_top_lines.py > gpurun_out/r2_ncu_${name}_source_top.txt 2>&1
  head -4 gpurun_out/r2_ncu_$name.txt | cut -c1-150; head -8 gpurun_out/r2_ncu_${name}_source_top.txt | cut -c1-200
}
cap k_inv3d_quad_dec 'k_inv3d<\(int\)1' 1 $B
cap k_inv3d_quad_scan 'k_inv3d<\(int\)2' 1 $B
echo "== config 2n"; SPERR_B200_VERBOSE=1 timeout 600 python scripts/bench_configs.py 2n > gpurun_out/r2u_bench_config2n.log 2>&1; grep bench_config gpurun_out/r2u_bench_config2n.log | cut -c1-1200; grep -v bench_config gpurun_out/r2u_bench_config2n.log | tail -4 | cut -c1-300
du -sh gpurun_out
