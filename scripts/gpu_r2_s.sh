#!/bin/bash
# round 2, call S: ncu evidence (launch list of one 1024^3 step, full captures of the hot kernels, summarised
# on the box: the reports themselves are too large to travel), other BASELINE configs
mkdir -p gpurun_out /tmp/rep
B="python bench.py --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 --check 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|DeviceRadixSort|DeviceScan|DeviceReduce' -c 1500 --csv --log-file gpurun_out/r2_launches_1024.csv $B > gpurun_out/r2_ncu_launches.log 2>&1; tail -1 gpurun_out/r2_ncu_launches.log | cut -c1-200
cap() {  # name, regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -c $cnt -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1
  python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_top_lines.py > gpurun_out/r2_ncu_${name}_source_top.txt 2>&1
  head -3 gpurun_out/r2_ncu_$name.txt | cut -c1-150
}
cap k_fwd3d_tma k_fwd3d_tma 1 $B
cap k_inv3d k_inv3d 1 $B
cap k_lis_plane k_lis_plane 1 $B
cap k_rec_apply k_rec_apply 1 $B
cap decode_1024 k_speck_decode_fast 2 $B
SPERR_B200_DEC_CLUSTER=8 cap decode_cluster8_512 k_speck_decode_fast 2 python scripts/dec_sweep.py 512 8
echo "== configs"; timeout 900 python scripts/bench_configs.py 3 4 5 2n > gpurun_out/r2_bench_configs.log 2>&1; grep bench_config gpurun_out/r2_bench_configs.log | cut -c1-700; tail -2 gpurun_out/r2_bench_configs.log | cut -c1-300
du -sh gpurun_out
