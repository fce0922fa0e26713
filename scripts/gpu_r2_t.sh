#!/bin/bash
# round 2, call T: the remaining BASELINE configs (4 on a box of whole chunks, 5 with the library's
# error text, noisy 2), full ncu captures of the level-0 inverse transforms and the reconstruction
# kernel with per-line stall samples
mkdir -p gpurun_out /tmp/rep
B="python bench.py --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 --check 0"
echo "== configs"; SPERR_B200_VERBOSE=1 timeout 900 python scripts/bench_configs.py 4 5 2n > gpurun_out/r2t_bench_configs.log 2>&1; grep bench_config gpurun_out/r2t_bench_configs.log | cut -c1-900; grep -v bench_config gpurun_out/r2t_bench_configs.log | tail -8 | cut -c1-300
cap() {  # name, regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -c $cnt -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1
  python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_top_lines.py > gpurun_out/r2_ncu_${name}_source_top.txt 2>&1
  head -4 gpurun_out/r2_ncu_$name.txt | cut -c1-150; head -12 gpurun_out/r2_ncu_${name}_source_top.txt | cut -c1-200
}
cap k_inv3d_dec 'k_inv3d<\(int\)1' 1 $B
cap k_inv3d_scan 'k_inv3d<\(int\)2' 1 $B
cap k_rec_apply 'k_rec_apply' 1 $B
cap k_fwd3d_tma 'k_fwd3d_tma' 1 $B
du -sh gpurun_out
echo "== step jitter"
Q="python bench.py --steps 8 --warmup 3 --e2e 0 --cpu-baseline 0 --check 0"
$Q 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', d['ms_per_step'], d['step_ms_each'], d['compress_ms_each'], d['decompress_ms_each'])"
$Q --diag noclocks 2>/dev/null | tail -1
$Q --diag noprof 2>/dev/null | tail -1
$Q --diag noclocks,noprof 2>/dev/null | tail -1
