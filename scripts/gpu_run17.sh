#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1; do
echo "== INV_PF=$pf"
SPERR_B200_INV_PF=$pf timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_rec_apply|k_inv3d' -c 20 --csv --log-file gpurun_out/l_$pf.csv python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > /dev/null 2>&1
grep -E "k_inv3d<[12]>|k_rec_apply" gpurun_out/l_$pf.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-60 | head -6
done
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-baseline 0 > gpurun_out/bench1024.log 2>&1; tail -1 gpurun_out/bench1024.log | cut -c1-1400
