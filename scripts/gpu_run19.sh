#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|DeviceRadixSort|DeviceScan|DeviceReduce' -c 900 --csv --log-file gpurun_out/launches1024.csv python bench.py --size 1024 --steps 1 --warmup 1 --e2e 0 --cpu-baseline 0 > gpurun_out/ncu_launches1024.log 2>&1
ls -la gpurun_out/launches1024.csv
