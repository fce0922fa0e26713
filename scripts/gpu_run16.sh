#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k 'regex:^k_rec|k_inv3d|k_fwd3d|k_lipref' -c 40 --csv --log-file gpurun_out/launches_rec.csv python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_rec.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_rec.csv') if not l.startswith('==')]
r=list(csv.reader(lines)); h=r[0]
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID'); ui=h.index('Metric Unit')
cur={}
for x in r[1:]:
    cur.setdefault((int(x[ii]),x[ki][:40]),{})[x[mi]]=(x[vi],x[ui])
for k in sorted(cur):
    print(k, {m:v for m,v in cur[k].items()})
PY
